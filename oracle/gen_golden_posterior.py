"""TEST INFRASTRUCTURE ONLY. Adds tests/golden/posterior_<case>.npz: the reference's accessory posterior outputs
(expectedTime1, expectedTime2_times / expectedTime2_aux / exp(likRep2), expectedGains; functions.cpp:582-757) on the
existing golden inputs, produced by the UNMODIFIED reference (oracle/_ref/libppm_ref.so).
Run:  python oracle/gen_golden_posterior.py      (needs /root/reference; the outputs are committed)."""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = [("ref_test", 2), ("syn12", 0), ("syn100", 0), ("cat40", 0)]  # (golden case, index of its parameter vector)

if __name__ == "__main__":
    for name, k in CASES:
        with open(os.path.join(GOLD, name + ".json")) as f:
            meta = json.load(f)
        p = np.array(meta["results"][k]["params"])
        R = ref.Reference(os.path.join(GOLD, name + ".nwk"), os.path.join(GOLD, name + ".pa.txt"), False)
        et1 = R.expected_time1(p)
        times, dist = R.expected_time2(p)
        gains = np.array([R.expected_gains(g) for g in (p[4], 0.5 * p[4], 1e-3)])
        np.savez_compressed(os.path.join(GOLD, "posterior_" + name + ".npz"), params=p, et1=et1, times=times, dist=dist,
                            gains_g2=np.array([p[4], 0.5 * p[4], 1e-3]), gains=gains)
        print(name, "nRep", R.nRep, "S", len(times), "et1[:3]", et1[:3], "gains", gains)
        R.close()
