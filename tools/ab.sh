#!/bin/bash
# same-box A/B: tools/ab.sh "<bench args>" variantA variantB ... ("default" = ntrtsim_b200/lib); 3 interleaved repetitions
args=$1; shift
for rep in 1 2 3; do
  for v in "$@"; do
    if [ "$v" = default ]; then unset TGB_LIB_DIR; else export TGB_LIB_DIR=$PWD/variants/$v; fi
    python bench.py --no-cpu-baseline --steps 5 $args 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v rep$rep value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))"
  done
done
