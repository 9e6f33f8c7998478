#!/bin/bash
# Round-end measurements on one GPU (outputs under gpurun_out/<TAG>_*): tests, smoke, default bench, reference arm, side benches.
TAG=$1
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; python tools/show_bench.py gpurun_out/${TAG}_bench_reference.json
python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; python tools/show_bench.py gpurun_out/${TAG}_bench_default.json
python bench.py --workload c5 --steps 5 > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; python tools/show_bench.py gpurun_out/${TAG}_bench_c5.json
python tests/perf_step.py > gpurun_out/${TAG}_bench_step.json 2> gpurun_out/${TAG}_bench_step.err; cat gpurun_out/${TAG}_bench_step.json
CGP_PREDICT_FUSED=1 python tests/perf_step.py > gpurun_out/${TAG}_bench_step_onepass.json 2> gpurun_out/${TAG}_bench_step_onepass.err; cat gpurun_out/${TAG}_bench_step_onepass.json
python tests/perf_facade.py > gpurun_out/${TAG}_bench_facade.json 2> gpurun_out/${TAG}_bench_facade.err; cat gpurun_out/${TAG}_bench_facade.json
python tools/bench_midsize.py > gpurun_out/${TAG}_bench_midsize.json 2> gpurun_out/${TAG}_bench_midsize.err; cat gpurun_out/${TAG}_bench_midsize.json
python tools/bench_large.py > gpurun_out/${TAG}_bench_large.json 2> gpurun_out/${TAG}_bench_large.err; cat gpurun_out/${TAG}_bench_large.json
python tests/perf_configs.py > gpurun_out/${TAG}_bench_configs.json 2> gpurun_out/${TAG}_bench_configs.err; cat gpurun_out/${TAG}_bench_configs.json
python tools/bench_2d_batch.py > gpurun_out/${TAG}_bench_2d.json 2> gpurun_out/${TAG}_bench_2d.err; cat gpurun_out/${TAG}_bench_2d.json
python tools/bench_perobject_fit.py > gpurun_out/${TAG}_perobject.json 2> gpurun_out/${TAG}_perobject.err; cat gpurun_out/${TAG}_perobject.json
