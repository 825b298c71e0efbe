"""TEST INFRASTRUCTURE ONLY.  Pure-Python/numpy restatement of the reference's pre-processing
(Inference::Inference, /root/reference/source/functions.cpp:76-156; RecTree, objects.cpp:38-149;
PAmatrix, objects.cpp:164-214, 254-281) producing the flat arrays oracle/ppm_oracle.c consumes,
(the PPM simulator used for seeded inputs lives in ppmmodelpangenome_b200/simulate.py).
The product's own host builder (ppmmodelpangenome_b200/csrc/host_model.cpp) is checked against this
and against dumps of the real reference (tests/golden/)."""
import numpy as np

from ppmmodelpangenome_b200.simulate import parse_newick  # independent (pure-Python) Newick reader


def read_matrix(path):
    """objects.cpp:164-214: header line ignored unless it starts with 'ccc' (then it carries the
    multiplicities); each further line 'genome,v,v,...'. Returns (names, M[genome][gene] uint8, counts)."""
    with open(path) as f:
        lines = f.read().split("\n")
    counts = None
    if lines[0][:3] == "ccc":
        counts = np.array([int(t) for t in lines[0].split(",")[1:]], np.int32)
    names, rows = [], []
    for line in lines[1:]:
        if not line:
            continue
        tok = line.split(",")
        names.append(tok[0])
        rows.append(np.array(tok[1:], dtype=np.uint8))
    M = np.stack(rows)
    if counts is None:
        counts = np.ones(M.shape[1], np.int32)
    return names, M, counts


def compress(P, counts):
    """objects.cpp:254-273: merge identical gene rows keeping first-occurrence order."""
    seen = {}
    keep, cnt, inverse = [], [], np.zeros(P.shape[0], np.int64)
    for i in range(P.shape[0]):
        k = P[i].tobytes()
        j = seen.get(k)
        if j is None:
            seen[k] = j = len(keep)
            keep.append(i); cnt.append(0)
        cnt[j] += int(counts[i])
        inverse[i] = j
    return P[keep], np.array(cnt, np.int32), inverse


def build_model(newick_text, names, M, counts):
    """Flat model. M is [genome][gene] as in the file; pattern r = gene column r."""
    left, right, bl, labels = parse_newick(newick_text)
    nn = len(left)
    nl = (nn + 1) // 2
    parent = np.full(nn, -1, np.int32)
    for v in range(nn):
        if left[v] >= 0:
            parent[left[v]] = v; parent[right[v]] = v
    # depths accumulated root -> leaf exactly as functions.cpp:66-75, then H - depth (:108-109)
    depth = np.zeros(nn)
    for v in range(nn):  # preorder: parent before child
        if left[v] >= 0:
            depth[left[v]] = depth[v] + bl[left[v]]
            depth[right[v]] = depth[v] + bl[right[v]]
    H = depth.max()
    height = H - depth
    # order by increasing height (:112-114); ties children-first (ids descend) -- see DESIGN.md
    order = np.array(sorted(range(nn), key=lambda v: (height[v], -v)), np.int32)
    # subtrees[rank] (:130-141)
    cur = [0]
    subs = []
    for i in range(nn - 1, 0, -1):
        v = int(order[i])
        cur = [x for x in cur if x != v]
        if left[v] >= 0:
            cur.append(int(left[v])); cur.append(int(right[v]))
        subs.append(list(cur))
    subs.reverse()
    sub_off = np.zeros(nn, np.int32)
    for r in range(nn - 1):
        sub_off[r + 1] = sub_off[r] + len(subs[r])
    sub_idx = np.array([x for s in subs for x in s], np.int32)
    # observations at leaf node ids (objects.cpp:364-371, 240-249)
    col = {nm: i for i, nm in enumerate(names)}
    nrep = M.shape[1]
    obs = np.zeros((nrep + 1, nn), np.int8)
    leaf_ids = [v for v in range(nn) if left[v] < 0]
    for v in leaf_ids:
        obs[:nrep, v] = M[col[labels[v]]]
    freq = M.sum(axis=0).astype(np.int32)
    # mrca (functions.cpp:46-65): last node in BFS order whose leaf set contains all present leaves;
    # with preorder ids that is the deepest common ancestor; an all-zero pattern gets the last BFS node.
    lo = np.full(nn, nn, np.int64); hi = np.full(nn, -1, np.int64)  # leaf-id range below each node
    for v in range(nn - 1, -1, -1):
        if left[v] < 0: lo[v] = hi[v] = v
        else:
            lo[v] = min(lo[left[v]], lo[right[v]]); hi[v] = max(hi[left[v]], hi[right[v]])
    bfs = [0]
    k = 0
    while k < len(bfs):
        v = bfs[k]; k += 1
        if left[v] >= 0: bfs += [int(left[v]), int(right[v])]
    mrca = np.zeros(nrep, np.int32)
    leaf_arr = np.array(leaf_ids)
    for r in range(nrep):
        pres = leaf_arr[obs[r, leaf_arr] == 1]
        if len(pres) == 0:
            mrca[r] = bfs[-1]; continue
        a, b = pres.min(), pres.max()
        v = 0
        while left[v] >= 0:
            if lo[left[v]] <= a and b <= hi[left[v]]: v = left[v]
            elif lo[right[v]] <= a and b <= hi[right[v]]: v = right[v]
            else: break
        mrca[r] = v
    # RecTree::treeLength (objects.cpp:115-122) in its own summation order
    tl = np.zeros(nn)
    for v in range(nn - 1, -1, -1):
        if left[v] >= 0:
            tl[v] = ((bl[left[v]] + tl[left[v]]) + bl[right[v]]) + tl[right[v]]
    L = float(tl[0])
    nobs = int(np.sum(counts))
    mean_genes = float(np.sum(freq.astype(np.int64) * counts) / nl)
    return dict(n_nodes=nn, n_leaves=nl, n_rep=nrep, n_obs=nobs, left=left, right=right, parent=parent, bl=bl,
                height=height, order=order, sub_off=sub_off, sub_idx=sub_idx, obs=obs,
                counts=np.asarray(counts, np.int32), mrca=mrca, freq=freq, H=float(H), L=L, mean_genes=mean_genes,
                labels=labels)


def model_from_files(tree_path, pa_path, dedupe=False):
    with open(tree_path) as f:
        text = f.readline()
    names, M, counts = read_matrix(pa_path)
    if dedupe:
        P, counts, _ = compress(np.ascontiguousarray(M.T), counts)
        M = np.ascontiguousarray(P.T)
    return build_model(text, names, M, counts)


