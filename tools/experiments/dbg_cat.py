import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cport, model
import ppmmodelpangenome_b200 as ppm
from ppmmodelpangenome_b200 import simulate
n, genes, seed = 10000, 24, 41000
rng = np.random.default_rng(seed)
nwk = simulate.sim_tree(n, rng, caterpillar=True)
names, M, truth = simulate.sim_genes(nwk, genes, rng)
open('/tmp/t.nwk', 'w').write(nwk + '\n'); simulate.write_matrix('/tmp/m.txt', names, M)
P = simulate.start_params(truth, genes)
orc = cport.Oracle(model.model_from_files('/tmp/t.nwk', '/tmp/m.txt'))
o = orc.eval(P, components=True)
inf = ppm.Inference('/tmp/t.nwk', '/tmp/m.txt', dedupe=True, device=0)
ll, i2 = inf.loglik_batch(P[None], return_i2=True)
print('ll gpu', ll[0], 'oracle', o['ll'], 'i2', i2[0], o['i2'])
c = inf.components(P, o['i2'])
oc = o['comp']
print('freq', inf.freq)
for j in range(inf.nRep):
    print(j, 'freq', inf.freq[j], 'mrca', inf.mrca[j], 'gpu', c[j], 'oracle lik0r lik1r lik1 lik2', oc[j])
