"""PPM simulator and small input helpers (pure Python / numpy).

Restates the reference's R simulation scripts (/root/reference/simulation_aux.R, run_simulations.R), which
cannot run here (no R): beta-splitting ultrametric trees with uniform coalescence depths (sim.tree,
simulation_aux.R:41-70) and gene presence/absence patterns of the three categories (sim.N :92-118,
sim.I1 :121-170, sim.I2 :173-238), written in the reference's input formats.  Used by bench.py and the
tests for seeded synthetic workloads; it is input generation, not part of the likelihood path."""
import math

import numpy as np


# ----------------------------------------------------------------------------- Newick
def parse_newick(text):
    """First line only (objects.cpp:123-131). Binary trees; every node carries ':length'; internal
    labels are skipped (objects.cpp:38-50). Returns preorder arrays (objects.cpp:104-114):
    left, right (-1 for leaves), bl (branch above node; root 0), labels ('' for internal)."""
    s = text.split("\n")[0].strip()
    if s.endswith(";"):
        s = s[:-1]
    left, right, bl, labels = [], [], [], []
    pos = 0

    def new_node():
        left.append(-1); right.append(-1); bl.append(0.0); labels.append("")
        return len(left) - 1

    # iterative recursive-descent: stack of (node, state)
    root = None
    stack = []  # nodes whose children are being read
    cur = None
    n = len(s)
    while pos < n:
        c = s[pos]
        if c == "(":
            v = new_node()
            if stack:
                p = stack[-1]
                if left[p] < 0: left[p] = v
                elif right[p] < 0: right[p] = v
                else: raise ValueError("non-binary node")
            else:
                root = v
            stack.append(v)
            pos += 1
        elif c == ",":
            pos += 1
        elif c == ")":
            cur = stack.pop()
            pos += 1
            # optional internal label, then optional :len
            j = pos
            while j < n and s[j] not in ":,()": j += 1
            pos = j
            if pos < n and s[pos] == ":":
                j = pos + 1
                while j < n and s[j] not in ",()": j += 1
                bl[cur] = float(s[pos + 1:j])
                pos = j
        else:
            j = pos
            while j < n and s[j] not in ":,()": j += 1
            v = new_node()
            labels[v] = s[pos:j]
            if stack:
                p = stack[-1]
                if left[p] < 0: left[p] = v
                elif right[p] < 0: right[p] = v
                else: raise ValueError("non-binary node")
            else:
                root = v
            pos = j
            if pos < n and s[pos] == ":":
                j = pos + 1
                while j < n and s[j] not in ",()": j += 1
                bl[v] = float(s[pos + 1:j])
                pos = j
    bl[0] = 0.0  # the reference never reads the root's own length (functions.cpp:95)
    return (np.array(left, np.int32), np.array(right, np.int32), np.array(bl, np.float64), labels)



# ----------------------------------------------------------------------------- simulator
def _fmt(x):
    return "%.10g" % x  # ape::write.tree writes 10 significant digits


def sim_tree(n, rng, height=1.0, b=-0.5, caterpillar=False):
    """simulation_aux.R:41-70: coalescence depths c(0, sort(U(0,height)^(n-2))), beta-splitting with
    p ~ Beta(b+1,b+1); caterpillar=True makes every split 1|rest. Returns a Newick string whose leaves
    are labelled 1..n. Iterative (no recursion limit)."""
    times = [0.0] + sorted(rng.uniform(0, height, n - 2).tolist()) if n > 2 else [0.0]
    if caterpillar and n > 200:
        # every split is 1 | rest: a chain; built directly (the general loop below is quadratic for it)
        perm = rng.permutation(n) + 1
        out = []
        for i in range(n - 1):
            t_par = times[i - 1] if i > 0 else 0.0
            out.append("(%d:%s," % (perm[i], _fmt(height - times[i])))
        out.append("%d:%s" % (perm[n - 1], _fmt(height - times[n - 2])))
        for i in range(n - 2, -1, -1):
            out.append("):%s" % _fmt(times[i] - (times[i - 1] if i > 0 else 0.0)))
        return "".join(out) + ";"
    # each work item: (t_parent, times(list, sorted), leaves(list)) -> builds string pieces
    # build an explicit tree first
    nodes = []  # (left, right, bl, label)

    def add(lft, rgt, blen, label):
        nodes.append([lft, rgt, blen, label]); return len(nodes) - 1

    root = add(-1, -1, 0.0, "")
    work = [(root, 0.0, times, list(range(1, n + 1)))]
    while work:
        v, t, tm, leaves = work.pop()
        k = len(leaves)
        if k == 1:
            nodes[v][2] = height - t; nodes[v][3] = str(leaves[0]); continue
        coal = tm[0]
        nodes[v][2] = coal - t
        rest = tm[1:]
        if k == 2:
            nl_ = 1
        elif caterpillar:
            nl_ = 1
        else:
            p = rng.beta(b + 1, b + 1)
            ks = np.arange(1, k)
            # log-binomial weights, stable for large k
            lw = (np.array([math.lgamma(k + 1) - math.lgamma(i + 1) - math.lgamma(k - i + 1) for i in ks])
                  + ks * math.log(max(p, 1e-300)) + (k - ks) * math.log(max(1 - p, 1e-300)))
            w = np.exp(lw - lw.max()); w /= w.sum()
            nl_ = int(rng.choice(ks, p=w))
        perm = rng.permutation(k)
        left_leaves = [leaves[i] for i in sorted(perm[:nl_])]
        right_leaves = [leaves[i] for i in sorted(perm[nl_:])]
        tperm = rng.permutation(len(rest))
        lt = sorted(rest[i] for i in tperm[:nl_ - 1])
        rt = sorted(rest[i] for i in tperm[nl_ - 1:])
        a = add(-1, -1, 0.0, ""); c = add(-1, -1, 0.0, "")
        nodes[v][0] = a; nodes[v][1] = c
        work.append((a, coal, lt, left_leaves)); work.append((c, coal, rt, right_leaves))
    # emit newick iteratively
    out = []
    st = [(root, 0)]
    while st:
        v, state = st.pop()
        lft, rgt, blen, label = nodes[v]
        if lft < 0:
            out.append(label + ":" + _fmt(blen)); continue
        if state == 0:
            out.append("("); st.append((v, 1)); st.append((lft, 0))
        elif state == 1:
            out.append(","); st.append((v, 2)); st.append((rgt, 0))
        else:
            out.append("):" + _fmt(blen))
    return "".join(out) + ";"


def sim_genes(newick_text, n_genes, rng, frac=(0.25, 0.35, 0.40), l0L=1.0, l1L=20.0, g2L=50.0, l2L=2500.0,
              e1=0.01, e2=0.01):
    """25 % Persistent (l0=1/L), 35 % Private (l1=20/L, arrival uniform on total branch
    length), 40 % Mobile (g2=50/L, l2=2500/L, arrival U(0,H)); observation errors e2 (1->0) and, for
    Mobile, e1 (0->1); all-zero genes are redrawn (simulation_aux.R:92-118,121-170,173-238).
    Vectorised over genes. Returns (labels-in-leaf-order names, M[genome][gene] uint8, truth dict)."""
    left, right, bl, labels = parse_newick(newick_text)
    nn = len(left)
    depth = np.zeros(nn)
    parent = np.full(nn, -1)
    for v in range(nn):
        if left[v] >= 0:
            for c in (left[v], right[v]):
                depth[c] = depth[v] + bl[c]; parent[c] = v
    H = depth.max(); L = bl[1:].sum()
    l0, l1, g2, l2 = l0L / L, l1L / L, g2L / L, l2L / L
    leaves = np.array([v for v in range(nn) if left[v] < 0])
    want = [int(round(n_genes * f)) for f in frac]
    want[2] = n_genes - want[0] - want[1]

    def two_state(state, t, g, l):  # exact CTMC transition over a branch, vectorised
        lam = g + l
        if lam == 0: return state
        pi1 = g / lam
        p1 = np.where(state == 1, pi1 + (1 - pi1) * np.exp(-lam * t), pi1 * (1 - np.exp(-lam * t)))
        return (rng.random(state.shape) < p1).astype(np.uint8)

    def draw(cat, m):
        st = np.zeros((nn, m), np.uint8)
        if cat == 0:
            st[0] = 1
            for v in range(1, nn):
                st[v] = st[parent[v]] & (rng.random(m) < np.exp(-l0 * bl[v]))
        elif cat == 1:
            # arrival: at the root with weight (1/l1) vs uniformly on branches with weight L
            at_root = rng.random(m) < (1 / l1) / (1 / l1 + L)
            edge = rng.choice(np.arange(1, nn), size=m, p=bl[1:] / L)
            frac_t = rng.random(m)
            st[0] = at_root
            for v in range(1, nn):
                inherit = st[parent[v]] & (rng.random(m) < np.exp(-l1 * bl[v]))
                born = (~at_root) & (edge == v) & (rng.random(m) < np.exp(-l1 * bl[v] * (1 - frac_t)))
                st[v] = inherit | born
        else:
            tarr = rng.uniform(0, H, m)  # depth of arrival, all lineages alive then start absent
            known = np.zeros((nn, m), bool)
            for v in range(1, nn):
                pv = parent[v]
                crosses = (depth[pv] <= tarr) & (tarr < depth[v])
                span = np.where(crosses, depth[v] - tarr, bl[v])
                src = np.where(crosses, 0, st[pv]).astype(np.uint8)
                act = crosses | known[pv]
                st[v] = np.where(act, two_state(src, span, g2, l2), 0)
                known[v] = act
        o = st[leaves]
        if cat == 2:
            flip = rng.random(o.shape)
            o = np.where(o == 1, flip >= e2, flip < e1).astype(np.uint8)
        else:
            o = (o & (rng.random(o.shape) >= e2)).astype(np.uint8)
        return o

    blocks = []
    for cat in range(3):
        got = np.zeros((len(leaves), 0), np.uint8)
        while got.shape[1] < want[cat]:
            o = draw(cat, max(64, 2 * (want[cat] - got.shape[1])))
            got = np.concatenate([got, o[:, o.sum(axis=0) > 0]], axis=1)
        blocks.append(got[:, :want[cat]])
    M = np.concatenate(blocks, axis=1)
    names = [labels[v] for v in leaves]
    truth = dict(N0=want[0], l0=l0, i1=want[1] * l1 / (1 + l1 * L), l1=l1, g2=g2, l2=l2, s_10=e2, s_01=e1, H=H, L=L)
    return names, M, truth


def write_matrix(path, names, M, counts=None):
    with open(path, "w") as f:
        if counts is None:
            f.write(",".join("V%d" % (i + 1) for i in range(M.shape[1])) + "\n")
        else:
            f.write("ccc," + ",".join(str(int(c)) for c in counts) + "\n")
        for nm, row in zip(names, M):
            f.write(nm + "," + ",".join(map(str, row.tolist())) + "\n")


def start_params(truth, nobs, rng=None, jitter=0.2):
    """Evaluation vectors = truth * exp(N(0, jitter^2)) clipped into the optim bounds
    (functions.cpp:395-396)."""
    p = np.array([truth["N0"], truth["l0"], max(truth["i1"], 1e-3), truth["l1"], truth["g2"], truth["l2"],
                  truth["s_10"], truth["s_01"]], float)
    if rng is not None:
        p = p * np.exp(rng.normal(0, jitter, 8))
    return p


def pack_patterns(newick_text, names, M):
    """Bit-pack M[genome][gene] into the C-ABI row format (include/ppm_b200.h): one row of ceil(n/32) uint32
    per gene, bit k = observation of the k-th leaf in preorder. Returns (left, right, bl, rows)."""
    left, right, bl, labels = parse_newick(newick_text)
    leaves = [v for v in range(len(left)) if left[v] < 0]
    row_of = {nm: i for i, nm in enumerate(names)}
    n = len(leaves)
    nw = (n + 31) // 32
    G = M.shape[1]
    rows = np.zeros((G, nw), np.uint32)
    for k, v in enumerate(leaves):
        bits = M[row_of[labels[v]]].astype(np.uint32)
        rows[:, k >> 5] |= bits << np.uint32(k & 31)
    return left, right, bl, rows
