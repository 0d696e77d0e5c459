mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1za.json 2> gpurun_out/bench_r1za.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1za.json')); print(d['value'], d['e2e']['value'], d['ms_per_step']); print(json.dumps(d.get('dxt_hc'))[:1500])"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:hc_tree_split_kernel -c 3 -f -o gpurun_out/hc_tree_split_r1z python tools/prof_hc.py 2048 --faces 6 --fmt DXT1 --reps 1 > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"hc_tiles_kernel|hc_palettize_kernel" -c 2 -f -o gpurun_out/hc_tiles_r1z python tools/prof_hc.py 2048 --faces 6 --fmt DXT1 --reps 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
