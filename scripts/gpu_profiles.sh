# GPU-box profile capture for a round: bench line, ncu launch list, ncu --set full of the headline kernels.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_profiles.sh'   then   python scripts/summarise_profiles.py r02
set -x
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
B="python bench.py --steps 1 --warmup 1 --no-solve --no-cpu-baseline --no-trt-like"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/bench_ncu.log 2>&1
tail -2 gpurun_out/bench_ncu.log
cap() { # name, kernel regex, skip, count
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -o gpurun_out/prof_$1 -f $B > gpurun_out/ncu_$1.log 2>&1; tail -1 gpurun_out/ncu_$1.log; }
cap conv 'conv64_halo_kernel' 2 2
cap dbscan 'db_scan' 4 4
cap conv_box64 'conv_umma_kernel<\(int\)64, \(bool\)1' 4 2
cap conv_mid 'conv_umma_kernel<\(int\)128' 16 8
cap conv_heads 'conv_umma_kernel<\(int\)80' 2 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:graph_solve -s 1 -c 1 -o gpurun_out/prof_solve -f python -c "
import sys; sys.path.insert(0, '.')
from omniswarm_b200 import host, synth
g = synth.pose_graph_c5(0); s = host.PoseGraphSolver(2048, 12288)
for _ in range(2): p, m = s.solve(g)
print(m.solve_ms, m.pcg_iterations)
" > gpurun_out/ncu_solve.log 2>&1
ls -la gpurun_out | tail -20
