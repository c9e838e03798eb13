#!/bin/bash
# Developer build: only the default DNA R = 4 variant of the lane = site kernel (EPA_DEV_MIN), ~40 s.
#   tools/devbuild.sh [extra nvcc flags]   ->  epa-ng_b200/libepa_dev.so   (use with EPA_B200_LIB=...)
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 -DEPA_DEV_MIN "$@" -shared -I include \
  -o epa-ng_b200/libepa_dev.so epa-ng_b200/csrc/epa_b200.cu $(ls epa-ng_b200/csrc/host/*.cpp | grep -v main.cpp) -lpthread
