for i in 1 2; do
for o in "fwd_smem_kb=101" "fwd_smem_kb=100"; do
  echo -n "$o: "; VCA_OPTS=$o python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"
done; done
