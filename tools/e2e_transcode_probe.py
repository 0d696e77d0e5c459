import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, numpy as np, crnsynth, crunch2_b200 as crn
ctx = crn.Context(0)
data = crnsynth.synth_crn(8192, 8192, "DXT5", seed=4, with_crc=False, n_color_ep=4096, n_color_sel=4096, n_alpha_ep=2048, n_alpha_sel=2048, skew=0.1)
tex = ctx.unpack_begin(data)
host_out = torch.empty(tex.total_size, dtype=torch.uint8).pin_memory().numpy()
d_out = torch.empty(tex.total_size, dtype=torch.uint8, device="cuda")
for rep in range(3):
    t0 = time.perf_counter(); t2 = ctx.unpack_begin(data); t1 = time.perf_counter()
    ctx._check(ctx._lib.crn_gpu_crnd_unpack_all_levels_host(t2._tex, host_out.ctypes.data, host_out.size)); t3 = time.perf_counter()
    t2.close(); t4 = time.perf_counter()
    tex.unpack_all_device(d_out, tex.total_size); t5 = time.perf_counter(); ctx.synchronize(); t6 = time.perf_counter()
    print("begin %.1f ms, unpack_all_host %.1f ms, close %.1f ms, device enqueue %.1f ms + sync %.1f ms" % ((t1-t0)*1e3, (t3-t1)*1e3, (t4-t3)*1e3, (t5-t4)*1e3, (t6-t5)*1e3))
