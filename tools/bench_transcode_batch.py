"""Batched CRN -> DXTn transcode (one crn_gpu_crnd_unpack_batch call over many files): CUDA events, L2 flushed, bit-exactness of a
sample against the oracle port.  Usage: python tools/bench_transcode_batch.py NFILES SIZE [FMT] [DISTINCT] [STEPS]
The files are DISTINCT streams cycled (neighbouring lanes never decode the same file), each with its own device slab."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import crnsynth  # noqa: E402
import helpers  # noqa: E402
import crunch2_b200 as crn  # noqa: E402


def run_batch(ctx, ext, dev, flush, nfiles, size, fmt="DXT5", distinct=64, steps=5, peak_gbs=6542.7, check=4, port=None):
    distinct = min(distinct, nfiles)
    datas = [crnsynth.synth_crn(size, size, fmt, seed=100 + i, with_crc=False, skew=0.1) for i in range(distinct)]
    texs = [ctx.unpack_begin(datas[i % distinct]) for i in range(nfiles)]
    per = texs[0].total_size
    nlev = texs[0].info["levels"]
    d_all = torch.empty(per * nfiles, dtype=torch.uint8, device=dev)
    ptrs = [d_all.data_ptr() + i * per for i in range(nfiles)]
    caps = [per] * nfiles
    l0 = ctx.launch_count
    ctx.unpack_batch(texs, ptrs, caps)
    launches = ctx.launch_count - l0
    bt = []
    for _ in range(steps):
        flush.fill_(4); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext); ctx.unpack_batch(texs, ptrs, caps); e1.record(ext); e1.synchronize()
        bt.append(e0.elapsed_time(e1))
    ms = sum(bt) / len(bt)
    texels = nfiles * sum(max(1, size >> l) ** 2 for l in range(nlev))
    in_bytes = sum(len(datas[i % distinct]) for i in range(nfiles))
    out = {"workload": "%d x crn_%s_%dx%d_%dlevels in one call (%d distinct streams cycled)" % (nfiles, fmt.lower(), size, size, nlev, distinct),
           "streams": nfiles * nlev, "value": texels / (ms / 1e3) / 1e9, "unit": "Gtexel/s", "ms": ms, "ms_min": min(bt), "launches_per_call": int(launches),
           "roofline": {"bound": "hbm", "achieved": (in_bytes + per * nfiles) / (ms / 1e3) / 1e9, "peak": peak_gbs, "unit": "GB/s",
                        "frac": (in_bytes + per * nfiles) / (ms / 1e3) / 1e9 / peak_gbs,
                        "algorithmic_bytes": int(in_bytes + per * nfiles)}}
    if port is not None and check:
        host = d_all.cpu().numpy()
        ok = True
        for i in list(range(min(check, nfiles))) + [nfiles - 1]:
            want = b"".join(b"".join(lv) for lv in helpers.port_unpack_all(port, datas[i % distinct]))
            ok = ok and host[i * per:(i + 1) * per].tobytes() == want
        out["bit_exact_vs_port_sample"] = bool(ok)
    for t in texs:
        t.close()
    return out


if __name__ == "__main__":
    nfiles = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    fmt = sys.argv[3] if len(sys.argv) > 3 else "DXT5"
    distinct = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
    ctx = crn.Context(0)
    dev = torch.device("cuda", 0)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    t0 = time.time()
    out = run_batch(ctx, ext, dev, flush, nfiles, size, fmt, distinct, steps, port=helpers.load_port())
    out["wall_s"] = time.time() - t0
    print(json.dumps(out))
