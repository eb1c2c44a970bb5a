"""Development aid: host-API (PCIe) pipeline of carta1_encode_pcm / carta1_decode_su, alone and in flight
together.  CARTA1_TRACE_PASSES=1 prints every pass's H2D / compute / D2H window."""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carta1_b200  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
upp = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(seconds * 44100)
frames = (n + 511) // 512
n_su = 2 * frames
g = torch.Generator().manual_seed(1)
pcm_h = (0.3 * torch.randn((2, n), generator=g)).pin_memory()
su_h = torch.empty(n_su * 212, dtype=torch.uint8).pin_memory()
su2_h = torch.empty(n_su * 212, dtype=torch.uint8).pin_memory()
out_h = torch.empty((2, frames * 512), dtype=torch.float32).pin_memory()
chans = [pcm_h[0].numpy(), pcm_h[1].numpy()]
outs = [out_h[0].numpy(), out_h[1].numpy()]
su, su2 = su_h.numpy(), su2_h.numpy()
opts = carta1_b200.make_enc_opts(fixed_block_modes=[0, 0, 0])
c1, c2 = carta1_b200.Context(0), carta1_b200.Context(0)
if upp:
    c1.set_max_units_per_pass(upp)
    c2.set_max_units_per_pass(upp)


def enc():
    c1.encode_pcm_into(chans, su, opts)


def dec():
    c2.decode_su_into(su2, n_su, 2, outs)


def duplex():
    th = threading.Thread(target=dec)
    th.start()
    enc()
    th.join()


def wall(fn, k=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3


enc()
su2[:] = su
trace = os.environ.pop("CARTA1_TRACE_PASSES", None)
te, td, tb = wall(enc), wall(dec), wall(duplex)
print("units/pass %d: encode %.2f ms  decode %.2f ms  sum %.2f ms (%.0f audio-s/s)  duplex %.2f ms (%.0f audio-s/s)" % (
    upp or 65536, te, td, te + td, seconds / (te + td) * 1e3, tb, seconds / tb * 1e3), flush=True)
if trace:
    os.environ["CARTA1_TRACE_PASSES"] = trace
    print("-- encode alone", file=sys.stderr); enc()
    print("-- decode alone", file=sys.stderr); dec()
    print("-- duplex", file=sys.stderr); duplex()

# pageable caller buffers (what a host that cannot pin its arrays passes): cudaMemcpyAsync stages through the driver
chans_p = [np.array(c) for c in chans]
outs_p = [np.empty_like(o) for o in outs]
su_p = np.empty_like(su)


def enc_p():
    c1.encode_pcm_into(chans_p, su_p, opts)


def dec_p():
    c2.decode_su_into(su_p, n_su, 2, outs_p)


tep, tdp = wall(enc_p), wall(dec_p)
print("pageable buffers: encode %.2f ms  decode %.2f ms  sum %.2f ms (%.0f audio-s/s)" % (
    tep, tdp, tep + tdp, seconds / (tep + tdp) * 1e3), flush=True)
t0 = time.perf_counter()
for a in chans_p + outs_p + [su_p]:
    torch.cuda.cudart().cudaHostRegister(a.ctypes.data, a.nbytes, 0)
t1 = time.perf_counter()
tep, tdp = wall(enc_p), wall(dec_p)
print("after cudaHostRegister (%.1f ms for %.2f GB): encode %.2f ms  decode %.2f ms" % (
    (t1 - t0) * 1e3, sum(a.nbytes for a in chans_p + outs_p + [su_p]) / 1e9, tep, tdp), flush=True)
for a in chans_p + outs_p + [su_p]:
    torch.cuda.cudart().cudaHostUnregister(a.ctypes.data)
