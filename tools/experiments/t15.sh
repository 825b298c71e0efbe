cd /root/repo
python - > gpurun_out/r2t_cli.txt 2>&1 <<'PY'
import subprocess, time, os, torch
G='tests/golden'
def run(tag, env={}):
    e=dict(os.environ); e.update(env)
    t0=time.time(); r=subprocess.run(['./ppmmodelpangenome_b200/inference','1',G+'/ref_test.nwk',G+'/ref_test.pa.txt','/tmp/m.txt','/tmp/c.txt'],capture_output=True,text=True,env=e); dt=time.time()-t0
    print(tag, 'rc', r.returncode, 'wall %.3f'%dt, flush=True)
run('no parent context')
x=torch.zeros(10, device='cuda'); torch.cuda.synchronize()
run('parent has a context')
big=torch.empty(40<<30, dtype=torch.uint8, device='cuda'); torch.cuda.synchronize()
run('parent holds 40 GB')
del big; torch.cuda.empty_cache()
run('parent freed')
import ppmmodelpangenome_b200 as ppm, numpy as np
from ppmmodelpangenome_b200 import simulate
rng=np.random.default_rng(1); nwk=simulate.sim_tree(1500,rng); names,M,truth=simulate.sim_genes(nwk,2000,rng)
left,right,bl,rows=simulate.pack_patterns(nwk,names,M)
inf=ppm.Inference.from_arrays(left,right,bl,rows,dedupe=True,device=0); P=np.stack([simulate.start_params(truth,2000)])
inf.loglik_batch(P); run('parent used band kernels')
inf.set_sparse(True); inf.loglik_batch(P); run('parent used sparse')
inf.close(); run('handle closed')
PY
