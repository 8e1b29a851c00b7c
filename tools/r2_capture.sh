# round-2 evidence run: GPU tests, bench (both arms), launch list, ncu --set full of poa_kernel and edlib_kernel (scratch driver script)
python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1; tail -2 gpurun_out/r2v_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; tail -2 gpurun_out/r2v_smoke.log
python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; tail -c 900 gpurun_out/r2v_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2v_bench_reference.json 2> gpurun_out/r2v_bench_reference.err; tail -c 600 gpurun_out/r2v_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-whole-program > gpurun_out/r2v_under_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'edlib_kernel' -c 1 -f -o gpurun_out/r2v_edlib_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-whole-program > gpurun_out/r2v_edlib_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'noisyreg_kernel' -s 1 -c 1 -f -o gpurun_out/r2v_noisyreg_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-whole-program > gpurun_out/r2v_noisyreg_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'poa_kernel' -c 1 -f -o gpurun_out/r2v_poa_full python tools/poa_prof.py 50 0 1000000 1 > gpurun_out/r2v_poa_ncu.log 2>&1
ls -la gpurun_out/r2v_*
