# static SASS instruction count of every kernel in libnixb200.so (hot loops must fit the 32 KB I-cache)
d=$(mktemp -d); cd $d; cuobjdump -xelf all ${1:-/root/repo/nix_b200/libnixb200.so} >/dev/null 2>&1
for f in *.sm_100a.cubin; do nvdisasm -c $f 2>/dev/null | awk -v f=$f '/^\.text\./{name=$0} /^ +\/\*[0-9a-f]+\*\//{n[name]++} END{for(k in n) printf "%6d  %s\n", n[k], substr(k,1,200)}'; done | sed -E 's/_ZN7nixb200[0-9]+_GLOBAL__N__[0-9a-f]+_[0-9]+_[a-z_]+_cu_[0-9a-f]+//; s/_ZN[0-9]+_INTERNAL_[0-9a-f]+_[0-9]+_[a-z_]+_cu_[0-9a-f]+//' | sort -rn | head -${2:-12}
rm -rf $d
