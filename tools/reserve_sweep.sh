# bench.py at several reservations: bash tools/reserve_sweep.sh TAG R1 R2 ...
tag=$1; shift
for r in "$@"; do LCD_BENCH_VERBOSE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --reserve-sms $r > gpurun_out/${tag}_bench_r$r.json 2> gpurun_out/${tag}_bench_r$r.err; tail -6 gpurun_out/${tag}_bench_r$r.err; python - <<PY
import json
j=json.load(open("gpurun_out/${tag}_bench_r$r.json"))
print("reserve $r:", round(j["value"],1), round(j["e2e"]["value"],1), round(j["ms_per_step"],1), round(j["ms_per_step_one_stream"],1), round(j["poa_ms_overlapped"],1), j["e2e_pileup_ms"])
PY
done
