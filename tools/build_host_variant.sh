# builds av1-go_b200/lib/libav1r_<tag>.so from the current CUDA objects and the host sources compiled with other flags
# usage: bash tools/build_host_variant.sh TAG "FLAGS"
TAG=$1; FLAGS=$2
CXX=${CXX:-g++}
mkdir -p build/obj_$TAG
for f in av1-go_b200/csrc/*.cpp; do
  $CXX $FLAGS -g -std=c++17 -fPIC -Wall -Wno-unused-function -Iinclude -I/usr/local/cuda/include -c $f -o build/obj_$TAG/$(basename $f .cpp).o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o av1-go_b200/lib/libav1r_$TAG.so build/obj_$TAG/*.o build/obj/kernels/*.cu.o build/obj/*.cu.o -lcudart -lpthread
ls -la av1-go_b200/lib/libav1r_$TAG.so
