#!/bin/bash
# builds build/variants/libqsmrt_<name>.so with extra nvcc flags (A/B experiments; QSMRT_LIB selects it at run time)
#   tools/build_variant.sh m0q10 "-DQSMRT_TRACE_MINB_M0Q=10"
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
d=build/variants/obj_$name
mkdir -p $d
cp pyqsm_b200/csrc/*.cu pyqsm_b200/csrc/*.cuh pyqsm_b200/csrc/*.h pyqsm_b200/csrc/Makefile $d/
mkdir -p build/variants/include && cp include/qsmrt.h build/variants/include/
sed -i 's#../../include#../include#g' $d/Makefile $d/*.cu $d/*.h $d/*.cuh 2>/dev/null || true
make -C $d -s EXTRA="$flags" OUT=../libqsmrt_$name.so
rm -rf $d
ls -la build/variants/libqsmrt_$name.so
