mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "elasticity or mesh or sampler" > gpurun_out/pytest_el.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_el.log
tail -5 gpurun_out/pytest_el.log
python tools/elastic_iter_list.py 3 > gpurun_out/el_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_elastic.csv python tools/elastic_iter_list.py 3 > gpurun_out/el_ncu.log 2>&1
tail -2 gpurun_out/el_ncu.log
