python -m pytest tests/test_gpu_splat.py tests/test_gpu_ewa.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -3
python bench_splat.py 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in d if k.startswith('ms_')}, d['kernels'])"
python bench_splat.py 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in d if k.startswith('ms_')}, d['kernels'])"
