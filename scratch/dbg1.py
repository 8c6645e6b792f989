import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import oracle as O, synth
from dspsr_b200 import engine as E
ctx = E.Context(0)
def case(input_real, input_nchan, npol, C, F, npos, nneg, npart, with_response=True, seed=1, max_npart=0):
    rng = np.random.default_rng(seed)
    nchan = input_nchan * C
    f = O.fb_sizes(input_real, input_nchan, npol, nchan, F, npos, nneg)
    ndim = 1 if input_real else 2
    ndat = npart * f.nsamp_step + f.nsamp_overlap
    x = rng.standard_normal((input_nchan, npol, ndat * ndim)).astype(np.float32)
    H = None
    if with_response:
        H = np.exp(1j * rng.uniform(-np.pi, np.pi, (nchan, F))).astype(np.complex64); H[0,0]=0
    ref = O.filterbank(f, x, H)
    eng = E.FilterbankEngine(ctx, input_real, input_nchan, npol, C, F, npos, nneg, H, max_npart)
    out = eng.perform(torch.from_numpy(x).cuda()).cpu().numpy().view(np.complex64)
    print("case", (input_real, input_nchan, npol, C, F, npos, nneg, npart, max_npart), "P,Q", eng.info.fft_rows, eng.info.fft_cols, "nkeep", f.nkeep)
    rms = np.sqrt((np.abs(ref)**2).mean())
    err = np.abs(out-ref)/rms
    e = err.reshape(nchan, npol, npart, f.nkeep)
    print("  err by part:", e.max(axis=(0,1,3)))
    print("  err by pol:", e.max(axis=(0,2,3)))
    ec = e.max(axis=(1,2,3)); print("  bad chans:", np.flatnonzero(ec>1e-5)[:20], "of", nchan)
    es = e.max(axis=(0,1,2)); bad = np.flatnonzero(es>1e-5); print("  bad samples:", bad[:10], "...", bad[-10:], len(bad), "of", f.nkeep)
case(1,1,2,16,64,5,6,11,max_npart=3)
case(1,1,2,16,64,5,6,11,max_npart=16)
case(1,1,2,16,64,5,6,3,max_npart=3)
case(0,2,2,32,1,0,0,20,with_response=False)
case(0,1,2,32,1,0,0,20,with_response=False)
case(1,1,2,2,8192,457,459,2)
case(1,1,2,256,8192,457,459,3)
