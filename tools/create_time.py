import time, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from cosmology_model_fit_b200 import Engine, datasets, fits
t0=time.perf_counter(); sn=datasets.pantheon_plus(cut=False); t1=time.perf_counter()
sp=fits.sn_pantheon(sn); t2=time.perf_counter()
e=Engine(sp); t3=time.perf_counter()
th=np.array([[-19.3,70.,0.3,0.]]); e.chi_squared(th); t4=time.perf_counter()
print(f"dataset {t1-t0:.2f}s  spec (cho_factor) {t2-t1:.2f}s  cl_create {t3-t2:.2f}s  first eval {t4-t3:.3f}s")
e.close()
