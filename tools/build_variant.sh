#!/bin/bash
# Build a variant of libagatha_b200.so for kernel A/B measurements: recompiles the packed-kernel translation units with extra
# -D flags and links them with the objects of the regular build (python -m agatha_b200.build first).
#   tools/build_variant.sh NAME "-DAGATHA_MB24=3 -DAGATHA_INLINE_EVENTS=1" [tu ...]     -> build/variants/NAME/libagatha_b200.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; FLAGS=$2; shift 2
TUS=${@:-extend16_inst_c24.cu extend16_inst_wide4.cu}
OUT=$ROOT/build/variants/$NAME
mkdir -p $OUT
OBJS=""
for f in $ROOT/agatha_b200/lib/*.o; do
  b=$(basename $f)
  skip=0
  for t in $TUS; do [ "$b" = "$t.o" ] && skip=1; done
  [ $skip = 0 ] && OBJS="$OBJS $f"
done
for t in $TUS; do
  /usr/local/cuda/bin/nvcc $FLAGS -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fopenmp --default-stream per-thread \
     -I$ROOT/include -I$ROOT/agatha_b200/csrc -c $ROOT/agatha_b200/csrc/$t -o $OUT/$t.o &
done
wait
for t in $TUS; do OBJS="$OBJS $OUT/$t.o"; done
/usr/local/cuda/bin/nvcc -shared -o $OUT/libagatha_b200.so $OBJS -Xcompiler -fopenmp -lgomp -lpthread
echo $OUT/libagatha_b200.so
