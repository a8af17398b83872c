O=gpurun_out/ev2
mkdir -p $O
python -m pytest tests -q -m gpu 2>&1 | tail -2 > $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1
python bench.py --steps 200 --warmup 10 > $O/r02_bench_default.json 2> $O/bench_default.err
for c in 3 4 5 1; do python bench.py --config $c --steps 100 --warmup 5 --no_also > $O/r02_bench_cfg$c.json 2> /dev/null; done
python scripts/epoch_timeline.py --packed 2>&1 | grep -v -i "warn" > $O/r02_epoch_timeline.txt
python scripts/step_timeline.py --config 2 --summary 2>&1 | grep -v -i "warn" > $O/r02_step_timeline_cfg2.txt
cat $O/pytest_gpu.txt $O/smoke.txt
