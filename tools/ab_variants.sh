# same-box A/B of library variants kept under build_variants/ (git-ignored): bash tools/ab_variants.sh base new ...
last=""
for v in "$@" "$@"; do
  cp build_variants/$v.so spvipes_b200/libspvipes_b200.so; last=$v
  echo "== $v"; timeout ${AB_TIMEOUT:-100} python bench.py ${AB_ARGS:---steps 1000 --warmup 50} --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])"
done
