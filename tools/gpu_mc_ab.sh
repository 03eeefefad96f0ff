B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-qformer-only --unfolded-steps 0 --parity-rows 4"
pick='import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(round(d["ms_per_step"],2),{k:round(v,2) for k,v in d["path"]["kernel_ms_per_step"].items()},d["clocks"]["sm_mhz"],d["parity_sample"]["ok"])'
for rep in 1 2; do
echo -n "MC=1: "; $B 2>/dev/null | python -c "$pick"
echo -n "MC=2: "; TDC_GEMM_MC=2 $B 2>/dev/null | python -c "$pick"
done
