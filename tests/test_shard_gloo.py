"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes partition the units of a batch without
overlap, and the timing reduction is a max over ranks (the contract bench.py --gpus N relies on)."""
import os
import socket
import subprocess
import sys

import helpers
from crunch2_b200 import shard

WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from crunch2_b200 import shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
dims = [(max(1, 2048 >> l), max(1, 2048 >> l)) for l in range(12)]
costs = shard.unit_costs(dims, faces=6)
mine = shard.partition_units(costs, world)[rank]
t = shard.max_over_ranks(10.0 + rank)                 # slowest rank defines the step time
total = shard.sum_over_ranks(sum(costs[i] for i in mine))
gathered = [None] * world
dist.all_gather_object(gathered, mine)
if rank == 0:
    print(json.dumps(dict(t=t, total=total, units=gathered, ncosts=len(costs), sumcosts=sum(costs))))
dist.destroy_process_group()
'''


def test_partition_is_balanced_and_complete():
    costs = shard.unit_costs([(max(1, 4096 >> l), max(1, 1024 >> l)) for l in range(13)], faces=1) * 5
    for world in (1, 2, 3, 8):
        parts = shard.partition_units(costs, world)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(costs)))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs)


def test_gloo_world_size_2(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), helpers.ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["t"] == 11.0
    assert d["total"] == d["sumcosts"]
    flat = sorted(i for u in d["units"] for i in u)
    assert flat == list(range(d["ncosts"]))


HC_WORKER = r'''
import os, sys, json, hashlib
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np
import torch, torch.distributed as dist
import blockgen, helpers, hc_util
import crunch2_b200 as crn
from crunch2_b200 import shard
from bench import mip_chain
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ctx = crn.Context(0, lib=helpers.load_sim())          # the real kernels under the SIMT emulator
out = {}
for fmt in (0, 3):
    img = blockgen.smooth_image(64, 48, 11, alpha=True)
    blocks, levels = hc_util.hc_layout([mip_chain(img)[:3]])
    whole = ctx.hc_compress(fmt, blocks, levels, codebook_sizes=(48, 48, 24, 48))
    part = ctx.hc_compress(fmt, blocks, levels, codebook_sizes=(48, 48, 24, 48), shard=(rank, world, shard.allgather_inplace))
    same = all(np.array_equal(whole[k], part[k]) for k in ("endpoint_indices", "selector_indices", "color_endpoints", "alpha_endpoints", "color_selectors", "alpha_selectors"))
    out[str(fmt)] = dict(same=bool(same), sha=hashlib.sha256(part["endpoint_indices"].tobytes() + part["selector_indices"].tobytes()).hexdigest())
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    print(json.dumps(gathered))
dist.destroy_process_group()
'''


def test_gloo_sharded_dxt_hc(tmp_path, sim):
    """One texture on two ranks: the per-cluster optimisation is dealt to the ranks, results all-gathered (gloo here, NCCL on
    the GPUs); every rank must end with exactly the unsharded result."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "hc_worker.py"
    script.write_text(HC_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), helpers.ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-3000:]
    import json
    ranks = json.loads([l for l in r.stdout.splitlines() if l.startswith("[")][-1])
    assert len(ranks) == 2
    for fmt in ("0", "3"):
        assert ranks[0][fmt]["same"] and ranks[1][fmt]["same"]
        assert ranks[0][fmt]["sha"] == ranks[1][fmt]["sha"]


CRN_WORKER = r'''
import os, sys, json, hashlib
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch, torch.distributed as dist
import blockgen, helpers
import crunch2_b200 as crn
from crunch2_b200 import shard
from bench import mip_chain
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ctx = crn.Context(0, lib=helpers.load_sim())
faces = [mip_chain(blockgen.smooth_image(64, 48, 11, alpha=True))[:3]]
whole, _, _ = ctx.compress_crn(faces, 2, quality_level=128)
part, rate, q = ctx.compress_crn(faces, 2, quality_level=128, shard=(rank, world, shard.allgather_inplace))
out = dict(same=bool(whole == part), sha=hashlib.sha256(part).hexdigest(), size=len(part))
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    print(json.dumps(gathered))
dist.destroy_process_group()
'''


def test_gloo_sharded_crn_compress(tmp_path, sim):
    """crn_compress to .CRN of one texture on two ranks (quantiser sharded by cluster, writer replicated): both ranks return the
    unsharded file, byte for byte."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "crn_worker.py"
    script.write_text(CRN_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), helpers.ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-3000:]
    import json
    ranks = json.loads([l for l in r.stdout.splitlines() if l.startswith("[")][-1])
    assert len(ranks) == 2 and ranks[0]["same"] and ranks[1]["same"]
    assert ranks[0]["sha"] == ranks[1]["sha"] and ranks[0]["size"] > 100


FAIL_WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch, torch.distributed as dist
import blockgen, helpers, hc_util
import crunch2_b200 as crn
from crunch2_b200 import shard
from bench import mip_chain
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
os.environ["CRN_B200_TEST_FAIL_RANK"] = "1"          # rank 1 fails inside the call, before the all-gather
ctx = crn.Context(0, lib=helpers.load_sim())
blocks, levels = hc_util.hc_layout([mip_chain(blockgen.smooth_image(64, 48, 11, alpha=True))[:2]])
msg = ""
try:
    ctx.hc_compress(0, blocks, levels, codebook_sizes=(48, 48, 24, 48), shard=(rank, world, shard.allgather_inplace))
except crn.CrnGpuError as e:
    msg = str(e)
gathered = [None] * world
dist.all_gather_object(gathered, msg)
if rank == 0:
    print(json.dumps(gathered))
dist.destroy_process_group()
'''


def test_gloo_sharded_failure_is_agreed_on(tmp_path, sim):
    """A rank that fails before the exchange must not leave the others waiting in the collective: it sends a status record on its way out
    (hc_host.h ShardAgreement), and every rank returns an error."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "fail_worker.py"
    script.write_text(FAIL_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), helpers.ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-3000:]
    import json
    msgs = json.loads([l for l in r.stdout.splitlines() if l.startswith("[")][-1])
    assert "injected" in msgs[1]
    assert "another rank failed" in msgs[0]
