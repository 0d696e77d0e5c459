"""Instruction mix of the hot kernels from the shipped library: `cuobjdump -sass` -> opcode histogram per kernel.
Usage: python tools/sass_histogram.py [lib.so] > profiles/sass_opcode_histogram_rN.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "crunch2_b200/libcrn_b200.so"
WANT = ["dxt1_optimize_clusters_kernel", "dxt1_optimize_clusters_cta_kernel", "dxt5_optimize_clusters_kernel", "pack_color_phase_kernel", "vq_fast_split_kernel",
        "vq_fast_tiny_kernel", "hc_tree_split_kernel", "transcode_tables_kernel", "transcode_walk_resolve_kernel", "transcode_levels_kernel", "crn_order_color_kernel",
        "unpack_blocks_kernel", "qdxt_training_kernel", "hc_tiles_kernel"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = name if any(w in name for w in WANT) else None
        if cur:
            hist[cur] = collections.Counter()
        continue
    if cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
        if m:
            hist[cur][m.group(1).split(".")[0]] += 1
print("cuobjdump -sass %s: static opcode counts (base mnemonic) of the hot kernels; sm_100a" % lib)
for name, h in hist.items():
    tot = sum(h.values())
    print("\n== %s  (%d instructions)" % (name[:160], tot))
    print("   " + "  ".join("%s %d" % (k, v) for k, v in h.most_common(22)))
    tensor = [k for k in h if k.startswith(("UTC", "HMMA", "IMMA", "UTMA", "TCGEN"))]
    print("   tensor / TMA opcodes: %s   (integer + byte work: none expected)" % (", ".join(tensor) or "none"))
