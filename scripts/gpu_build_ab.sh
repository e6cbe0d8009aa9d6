# A/B of graph-build variants in one gpurun call.  usage: gpu_build_ab.sh "N D R L" name ...   (name = default or a variants/lib_<name>.so)
shape="$1"; shift
for name in "$@"; do
  libenv=""; [ "$name" != "default" ] && libenv="$PWD/diskrag_b200/variants/lib_$name.so"
  echo "== $name"; DISKRAG_B200_LIB=$libenv timeout 600 python tests/tools/build_ab.py $shape 2>&1 | grep "BUILD_AB\|Error\|error" | tail -3
done
