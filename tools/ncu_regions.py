"""Group the SASS of one kernel (ncu `--page source --csv` export) into regions of equal execution count.

usage: python tools/ncu_regions.py source.csv [n_warp_items] [--sass]

Consecutive instructions executed (almost) equally often belong to one loop body; per region the tool prints the
instruction count, how many are FP64 (DFMA/DMUL/DADD/DSETP), executions per warp item, the share of all executed
instructions, the share of warp-time samples, and the stall reasons that dominate.  This is how the per-loop
budgets in profiles/*_regions.txt are produced.
"""
import csv
import sys


def main():
    path = sys.argv[1]
    items = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 0.0
    show_sass = "--sass" in sys.argv
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    ins = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            ex = float(r[col["Instructions Executed"]] or 0)
        except ValueError:
            continue
        smp = float(r[col["# Samples"]] or 0)
        st = {h: float(r[col[h]] or 0) for h in stall_cols}
        ins.append((r[col["Address"]], r[col["Source"]], ex, smp, st))
    tot_ex = sum(i[2] for i in ins)
    tot_smp = sum(i[3] for i in ins)

    def is_fp64(s):
        op = s.split()[0] if not s.startswith("@") else s.split()[1]
        return op.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")

    tot_f = sum(i[2] for i in ins if is_fp64(i[1]))
    if items:
        print(f"# instr/item {tot_ex / items:.1f}  fp64/item {tot_f / items:.1f}  samples {tot_smp:.0f}")
    else:
        print(f"# instr {tot_ex:.0f} fp64 {tot_f:.0f} samples {tot_smp:.0f}")
    # regions: consecutive instructions whose execution count is within 2 %
    regs = []
    cur = []
    for i in ins:
        if cur and (abs(i[2] - cur[-1][2]) > 0.02 * max(i[2], cur[-1][2], 1)):
            regs.append(cur)
            cur = []
        cur.append(i)
    if cur:
        regs.append(cur)
    for g in regs:
        ex = sum(i[2] for i in g)
        if ex < 0.003 * tot_ex:
            continue
        smp = sum(i[3] for i in g)
        nf = sum(1 for i in g if is_fp64(i[1]))
        fx = sum(i[2] for i in g if is_fp64(i[1]))
        st = {}
        for i in g:
            for h, v in i[4].items():
                st[h] = st.get(h, 0) + v
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        tops = " ".join(f"{h[6:]}={v / max(smp, 1):.2f}" for h, v in top)
        x = g[0][2] / items if items else g[0][2]
        print(f"off {g[0][0][-5:]} n {len(g):4d} fp64 {nf:3d} x {x:8.1f} exec {100 * ex / tot_ex:5.1f}% samples {100 * smp / tot_smp:5.1f}% "
              f"fp64share {100 * fx / max(tot_f, 1):5.1f}% rel {(smp / tot_smp) / (ex / tot_ex):.2f}  {tops}")
        if show_sass:
            for i in g:
                print(f"      {i[0][-5:]} {i[2] / items if items else i[2]:8.1f} {i[3]:7.0f}  {i[1]}")


if __name__ == "__main__":
    main()
