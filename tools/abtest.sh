# usage: bash tools/abtest.sh "<bench args>" <lib suffixes...>   (experiment builds of nix_b200.build --out=libnixb200_<sfx>.so)
args="$1"; shift
for v in "$@"; do
  lib=$PWD/nix_b200/libnixb200$v.so
  NIXB200_LIB=$lib python bench.py $args --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('variant[$v]', round(d['value']/1e9,3), {k:round(x,2) for k,x in d['phases_ms_per_step'].items()})
    elif l.strip(): print(l.strip()[:200])
"
done
