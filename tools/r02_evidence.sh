set -x
mkdir -p gpurun_out
SAN_TESTS="tests/test_parity_gpu.py tests/test_round2_gpu.py tests/test_tracking.py"
SAN_K="golden or ragged or dist_only or custom_mu or sweep_select_matches or binned or strided or peer_comm or sharded_helper or backward_matches or tracker"
timeout 1500 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 9 --log-file gpurun_out/r02_sanitizer_memcheck.log python -m pytest $SAN_TESTS -m gpu -x -q -k "$SAN_K" > gpurun_out/r02_sanitizer_memcheck_pytest.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck_pytest.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r02_sanitizer_racecheck.log python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -x -q -k "golden or ragged or sweep_select_matches or binned_walk and not 70001 or peer_comm" > gpurun_out/r02_sanitizer_racecheck_pytest.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck_pytest.txt
for f in gpurun_out/r02_sanitizer_memcheck_pytest.txt gpurun_out/r02_sanitizer_racecheck_pytest.txt gpurun_out/r02_sanitizer_memcheck.log gpurun_out/r02_sanitizer_racecheck.log; do tail -n 3 $f; done
for c in ${CASES:-grid visible sweep backward pca dist16m mask cfg5}; do   # ~7 MB of .ncu-rep each; gpurun_out/ is capped at 64 MiB per call
  case $c in sweep) K=field_sweep;; backward) K=field_backward;; pca) K="pca_project|field_tile";; *) K=field_tile;; esac
  ncu --set full --clock-control none --import-source on -k "regex:$K" --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_$c python tools/profile_case.py $c 2>&1 | tail -1
done
D3F_BENCH_LOAD_STEPS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-extras --no-cpu --no-e2e > gpurun_out/r02_launches_bench.json 2>&1
tail -c 300 gpurun_out/r02_launches_bench.json
