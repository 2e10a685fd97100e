# Everything the round-2 tables in DESIGN.md / profiles/ were filled from, in one gpurun call (one B200):
#   gpurun --timeout 1500 -- 'bash scripts/final_sweep.sh'   -> gpurun_out/f1/
mkdir -p gpurun_out/f1
O=gpurun_out/f1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:gemm_chain_kernel -s 6 -c 3 -o $O/chain python scripts/bench_chain.py > $O/ncu_chain.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "Warning\|_maybe_warn" | tail -12 > $O/tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
python bench.py > $O/bench_resnet50.json 2> $O/bench_resnet50.err
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-context"
$B --model resnet18 > $O/bench_resnet18.json 2>/dev/null
$B --model resnet152 > $O/bench_resnet152.json 2>/dev/null
$B --model lenet5 > $O/bench_lenet5.json 2>/dev/null
$B --precision bf16x3 > $O/bench_bf16x3.json 2>/dev/null
$B --precision bf16x3 --layout nchw > $O/bench_bf16x3_nchw.json 2>/dev/null
$B --layout nchw > $O/bench_bf16_nchw.json 2>/dev/null
python bench.py --mode invert --model resnet50 --steps 10 --warmup 3 > $O/invert_resnet50.json 2>/dev/null
python bench.py --mode invert --model resnet152 --steps 10 --warmup 3 > $O/invert_resnet152.json 2>/dev/null
python scripts/bench_chain.py > $O/chain.json 2>/dev/null
python scripts/bench_k3k5.py > $O/k3k5.json 2>/dev/null
python scripts/chain_timeline.py > $O/chain_timeline.txt 2>&1
python scripts/step_trace.py > $O/step_trace.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base function -k regex:"syrk|cast_bf16|round_tf32|split_bf16|pack_smallc|nchw_to_nhwc" -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-context > $O/launches_run.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2>/dev/null
tail -5 $O/tests.log; cat $O/smoke.log | tail -2; cat $O/bench_resnet50.json | head -c 600; echo; ls -la $O
