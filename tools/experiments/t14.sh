cd /root/repo
python - > gpurun_out/r2s_cli.txt 2>&1 <<'PY'
import subprocess, time, os
G='tests/golden'
for env in ({}, {}, {'CUDA_MODULE_LOADING':'EAGER'}, {'PPM_VERBOSE':'1'}):
    e=dict(os.environ); e.update(env)
    t0=time.time(); r=subprocess.run(['./ppmmodelpangenome_b200/inference','1',G+'/ref_test.nwk',G+'/ref_test.pa.txt','/tmp/m.txt','/tmp/c.txt'],capture_output=True,text=True,env=e); dt=time.time()-t0
    print(env, 'rc', r.returncode, 'wall %.3f'%dt, r.stderr[-300:].replace('\n',' | '))
PY
