#!/bin/bash
# Round-end evidence on one B200 (run through gpurun): tests, bench lines of every config, timelines, ncu launch list and one
# `ncu --set full` capture of a captured cfg2 step.  Everything lands in gpurun_out/ev/ (copy what is to be kept to profiles/).
O=gpurun_out/ev
mkdir -p $O
python -m pytest tests -q -m gpu 2>&1 | tail -3 > $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1
python bench.py --steps 200 --warmup 10 > $O/r02_bench_default.json 2> $O/bench_default.err
for c in 1 3 4 5; do python bench.py --config $c --steps 100 --warmup 5 --no_also > $O/r02_bench_cfg$c.json 2> /dev/null; done
python bench.py --impl reference --steps 3 --warmup 3 > $O/r02_bench_reference_arm.json 2> /dev/null
for c in 2 3 4 5; do python scripts/step_timeline.py --config $c --summary 2>&1 | grep -v -i "warn" > $O/r02_step_timeline_cfg$c.txt; done
for c in 2 3 5; do python scripts/gemm_launch_table.py --config $c; done > $O/r02_gemm_launch_table.txt 2>&1
python scripts/epoch_timeline.py 2>&1 | grep -v -i "warn" > $O/r02_epoch_timeline.txt
# ncu launch list of the bench command (the same command has just exited 0 without ncu)
if python bench.py --steps 3 --warmup 3 --no_cpu_baseline --no_also > /dev/null 2>&1; then
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
      python bench.py --steps 3 --warmup 3 --no_cpu_baseline --no_also > $O/ncu_launch.log 2>&1
  python scripts/summarize_launches.py $O/launches.csv > $O/r02_launches_bench_summary.txt
  gzip -f $O/launches.csv
fi
# one captured cfg2 step under ncu --set full (report stays on the box; only the text extract comes back)
if python scripts/ncu_step.py --config 2 --steps 1 > /dev/null 2>&1; then
  ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/step2 \
      python scripts/ncu_step.py --config 2 --steps 1 > $O/ncu_full.log 2>&1
  python scripts/ncu_extract.py /tmp/step2.ncu-rep > $O/r02_ncu_step_cfg2.txt 2>&1
fi
ls -la $O
