#!/bin/bash
# build_variant.sh NAME "-DFLAG=..."  ->  build/variants/libspcu_NAME.so : the library with trace_kernels.cu (and, with a third
# argument "shade", shade_kernels.cu) recompiled under extra flags, every other object taken from the regular build.
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
NAME=$1; FLAGS=$2; WHICH=${3:-trace}
OBJ=$ROOT/build/obj/csrc; OUT=$ROOT/build/variants; mkdir -p $OUT/$NAME
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -I$ROOT/include -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v"
OBJS=""
for o in comm trace_kernels shade_kernels path_kernels smwave_kernels spcu_api spcu_render build_kernels image_kernels mesh_kernels; do
  if [ "$o" = "trace_kernels" ] && [[ "$WHICH" == *trace* ]]; then
    $NV --fmad=false $FLAGS -c $ROOT/simplepath_b200/csrc/trace_kernels.cu -o $OUT/$NAME/trace_kernels.o 2> $OUT/$NAME/trace_kernels.ptxas.log
    OBJS="$OBJS $OUT/$NAME/trace_kernels.o"
  elif [ "$o" = "smwave_kernels" ] && [[ "$WHICH" == *smwave* ]]; then
    $NV --use_fast_math -ftz=false $FLAGS -c $ROOT/simplepath_b200/csrc/smwave_kernels.cu -o $OUT/$NAME/smwave_kernels.o 2> $OUT/$NAME/smwave_kernels.ptxas.log
    OBJS="$OBJS $OUT/$NAME/smwave_kernels.o"
  elif [ "$o" = "shade_kernels" ] && [[ "$WHICH" == *shade* ]]; then
    $NV --use_fast_math $FLAGS -c $ROOT/simplepath_b200/csrc/shade_kernels.cu -o $OUT/$NAME/shade_kernels.o 2> $OUT/$NAME/shade_kernels.ptxas.log
    OBJS="$OBJS $OUT/$NAME/shade_kernels.o"
  else
    OBJS="$OBJS $OBJ/$o.o"
  fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libspcu_$NAME.so $OBJS -ldl
echo built $OUT/libspcu_$NAME.so
