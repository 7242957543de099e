import os, sys, time, gzip
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import synth
from test_gpu_ingest import synth_problem, engine_for, fastq_text
t0=time.perf_counter()
import torch
torch.zeros(1, device='cuda'); torch.cuda.synchronize()
print('cuda init', time.perf_counter()-t0)
cfg = synth.make_config("C2", n_reads=int(os.environ.get("PROBE_READS", "60000")), seed=1)
b = cfg.generate(0, cfg.n_reads)
reads = cfg.reads_as_strings(b); f0, f1 = cfg.fields_as_strings(b)
text = fastq_text(reads, f0, f1, odd=False)
p='/dev/shm/probe.fastq.gz'
open(p,'wb').write(gzip.compress(text,1))
pin = synth_problem(cfg)
os.environ['SCAR_INGEST_TRACE']='1'
for it in range(3):
    t0=time.perf_counter()
    eng = engine_for(pin, 160, table_capacity=1<<18)
    t1=time.perf_counter()
    eng.set_verify(True)
    run = eng.process_fastq_file(p)
    t2=time.perf_counter()
    c=eng.counters(); ph=eng.phase_hist(); tab=eng.table()
    t3=time.perf_counter()
    eng.close()
    t4=time.perf_counter()
    print('iter',it,'ctx create %.3f  file %.3f  results %.3f  close %.3f'%(t1-t0,t2-t1,t3-t2,t4-t3), run.n, len(tab))
os.remove(p)
