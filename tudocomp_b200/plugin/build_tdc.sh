#!/usr/bin/env bash
# Builds the reference's `tdc` driver offline in three flavours (see registry_gpu.py):
#   build/tdc_ref        unmodified reference subset (CPU)            — byte-identity checker
#   build/tdc_gpu        reference registry + GpuTextDS ("mixed")     — -a "lzss_lcp(huff, gpu)"
#   build/tdc_gpu_only   GPU text index as the default ("only")       — same -a strings as the reference
#   build/tdc_block_ref / build/tdc_block_gpu   block mode driver (tdc_block.cpp) over the reference / GPU-only registry
# The reference sources are compiled where they lie under $REF (nothing is copied); the two missing third-party headers
# come from tudocomp_b200/plugin/shim.  Needs $REF, so it only runs in the build container; the binaries travel in build/.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
REF="${REF:-/root/reference}"
OUT="$ROOT/build/tdc"
JOBS="${JOBS:-8}"
[ -d "$REF/include/tudocomp" ] || { echo "build_tdc.sh: $REF absent, keeping prebuilt binaries"; exit 0; }
[ -f "$ROOT/tudocomp_b200/libtdcgpu.so" ] || make -s -C "$ROOT/tudocomp_b200/csrc"
CXXFLAGS="-std=gnu++14 -O2 -DNDEBUG -w -I$ROOT/tudocomp_b200/plugin/shim -I$ROOT/tudocomp_b200/plugin/include -I$ROOT/include -I$REF/include"
for mode in none mixed only; do
  case $mode in none) bin=tdc_ref;; mixed) bin=tdc_gpu;; only) bin=tdc_gpu_only;; esac
  gen="$OUT/gen_$mode"; rm -rf "$gen"; mkdir -p "$gen/tudocomp"
  # generated headers the reference's CMake would configure
  printf '#pragma once\n' > "$gen/tudocomp/config.h"
  printf '#pragma once\n#include <string>\nnamespace tdc { const std::string VERSION = "tdc-b200 offline build (%s)"; }\n' "$mode" > "$gen/tudocomp/version.hpp"
  TDC_GPU_MODE=$mode python3 "$REF/etc/genregistry.py" "$ROOT/tudocomp_b200/plugin/registry_gpu.py" "$gen/tudocomp/config.h" "$gen" --generate_files >/dev/null
  extra=""; [ "$mode" = only ] && extra="-DTDC_GPU_DEFAULT_TEXTDS"
  srcs=$(ls "$gen"/*.cpp)
  all="$srcs $REF/src/tudocomp_driver/tudocomp_driver.cpp $REF/src/tudocomp_stat/StatPhase.cpp $REF/src/tudocomp_stat/malloc.cpp"
  objs=""
  for s in $all; do objs="$objs $gen/$(basename "${s%.cpp}").o"; done
  for s in $all; do
    echo "g++ $CXXFLAGS $extra -I$gen -c $s -o $gen/$(basename "${s%.cpp}").o"
  done | xargs -P "$JOBS" -I{} sh -c '{}'
  if [ "$mode" != none ]; then
    g++ -o "$ROOT/build/$bin" $objs -L"$ROOT/tudocomp_b200" -ltdcgpu '-Wl,-rpath,$ORIGIN/../tudocomp_b200' -ldl
  else
    g++ -o "$ROOT/build/$bin" $objs -ldl
  fi
  echo "built build/$bin"
  # block mode driver (tdc_block.cpp): the same registry objects with its own main instead of tudocomp_driver.cpp
  if [ "$mode" != mixed ]; then
    [ "$mode" = none ] && bbin=tdc_block_ref || bbin=tdc_block_gpu
    g++ $CXXFLAGS $extra -I$gen -c "$ROOT/tudocomp_b200/plugin/tdc_block.cpp" -o "$gen/tdc_block.o"
    bobjs=$(echo $objs | tr ' ' '\n' | grep -v tudocomp_driver.o | tr '\n' ' ')
    if [ "$mode" = only ]; then
      # tdc_block_gpu runs one worker THREAD per GPU (CUDA starts once per process), so its registry units are a second
      # compile of the same generated sources with the reference's own -DSTATS_DISABLED (StatPhase is not thread-safe)
      ns="$OUT/gen_only_nostats"; rm -rf "$ns"; mkdir -p "$ns"
      nsobjs=""
      for s in $srcs $REF/src/tudocomp_stat/StatPhase.cpp; do nsobjs="$nsobjs $ns/$(basename "${s%.cpp}").o"; done
      for s in $srcs $REF/src/tudocomp_stat/StatPhase.cpp; do
        echo "g++ $CXXFLAGS $extra -DSTATS_DISABLED -I$gen -c $s -o $ns/$(basename "${s%.cpp}").o"
      done | xargs -P "$JOBS" -I{} sh -c '{}'
      g++ $CXXFLAGS $extra -DSTATS_DISABLED -DTDC_BLOCK_THREADS -I$gen -c "$ROOT/tudocomp_b200/plugin/tdc_block.cpp" -o "$ns/tdc_block.o"
      g++ -o "$ROOT/build/$bbin" $nsobjs "$ns/tdc_block.o" -L"$ROOT/tudocomp_b200" -ltdcgpu '-Wl,-rpath,$ORIGIN/../tudocomp_b200' -ldl -lpthread
    else
      g++ -o "$ROOT/build/$bbin" $bobjs "$gen/tdc_block.o" -ldl
    fi
    echo "built build/$bbin"
  fi
done
# end-to-end harness over the plugin classes (tdc_plugin_bench.cpp): single translation unit, no registry needed
gen="$OUT/gen_only"
g++ $CXXFLAGS -I"$gen" "$ROOT/tudocomp_b200/plugin/tdc_plugin_bench.cpp" "$REF/src/tudocomp_stat/StatPhase.cpp" -o "$ROOT/build/tdc_plugin_bench" \
  -L"$ROOT/tudocomp_b200" -ltdcgpu '-Wl,-rpath,$ORIGIN/../tudocomp_b200' -ldl
echo "built build/tdc_plugin_bench"
# the same harness against a WIDE-INDEX build of the reference (-DLEN_BITS=40: 64-bit len_t, def.hpp:100-114): shows that the
# plugin compiles there and checks its archives against the reference's own wide-index code (tests/test_plugin.py)
g++ $CXXFLAGS -DLEN_BITS=40 -I"$gen" "$ROOT/tudocomp_b200/plugin/tdc_plugin_bench.cpp" "$REF/src/tudocomp_stat/StatPhase.cpp" -o "$ROOT/build/tdc_plugin_bench40" \
  -L"$ROOT/tudocomp_b200" -ltdcgpu '-Wl,-rpath,$ORIGIN/../tudocomp_b200' -ldl
echo "built build/tdc_plugin_bench40"
# checker of the per-provider GPU classes (GpuProviders.hpp): TextDS<GpuSA, GpuPhi, GpuPLCP, GpuLCP, GpuISA> against TextDS<>
g++ $CXXFLAGS -I"$gen" "$ROOT/tudocomp_b200/plugin/tdc_providers_check.cpp" "$REF/src/tudocomp_stat/StatPhase.cpp" -o "$ROOT/build/tdc_providers_check" \
  -L"$ROOT/tudocomp_b200" -ltdcgpu '-Wl,-rpath,$ORIGIN/../tudocomp_b200' -ldl
echo "built build/tdc_providers_check"

