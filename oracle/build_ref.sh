#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference (kmtricks v1.6.0) CPU
# pipeline from the sources where they lie under /root/reference, with plain g++/gcc
# (the reference's own cmake build system is NOT run).  Outputs go only to oracle/_ref/
# (git-ignored, shipped to the GPU box by gpurun).  Nothing from the reference is copied
# into the repository: the few cmake-"configure_file" headers the sources expect
# (compile-time constants such as the k-mer size list) are re-created here with the
# values the reference's CMakeLists would have substituted (CMakeLists.txt:30-41,112;
# thirdparty/gatb-core-stripped/CMakeLists.txt "configure_file" block).
#
# Usage: oracle/build_ref.sh [REF=/root/reference] [JOBS=nproc] [MARCH=] [SUFFIX=]   (env vars)
# SUFFIX=_v3 MARCH=-march=x86-64-v3 builds a second CLI binary bin/kmtricks_v3 (AVX2/BMI2/FMA code generation, what
# the reference's -DNATIVE=ON gives on the bench hosts) for the timed CPU baseline; bench.py picks it when the host
# CPU has those flags.  The default (portable) build is the parity oracle and also builds the harness and the plugins.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
JOBS=${JOBS:-$(nproc)}
MARCH=${MARCH:-}            # e.g. MARCH=-march=native for a timed baseline on the same box
SUFFIX=${SUFFIX:-}
[ -d "$REF/include/kmtricks" ] || { echo "reference not found at $REF" >&2; exit 3; }
GEN=$OUT/gen; OBJ=$OUT/obj$SUFFIX
mkdir -p "$GEN/include/kmtricks" "$GEN/include/gatb/system/api" "$GEN/src/gatb/template" "$GEN/kff" "$OBJ" "$OUT/bin"
T=$REF/thirdparty; G=$T/gatb-core-stripped

# ---- configure_file substitutes (constants only) ---------------------------------
cat > "$GEN/include/kmtricks/config.hpp" <<'EOF'
#pragma once
#define KMER_LIST 32,64
#define KMER_LIST_STR "32,64"
#define KMER_N 2
#define PROJECT_NAME "kmtricks"
#define PROJECT_VER  "v1.6.0"
#define PROJECT_VER_MAJOR "1"
#define PROJECT_VER_MINOR "6"
#define PROJECT_VER_PATCH "0"
#define PROJECT_DESC "kmtricks - k-mer matrices and Bloom filters construction."
#define CONTACT "teo.lemane@genoscope.cns.fr"
#define HOST_SYSTEM "Linux"
#define COMPILER_C "GNU"
#define COMPILER_CXX "GNU"
#define CONDA_BUILD "OFF"
#define STATIC_BUILD "OFF"
#define NATIVE_BUILD "OFF"
#ifndef ARCH_X86_64
#define ARCH_X86_64
#endif
#define ARCH_FLAGS ""
#define MODULES_BUILD "ON"
#define DEV_BUILD "OFF"
#define GIT_SHA1 ""
#define BCLI_SHA1 ""
#define FMT_SHA1 ""
#define KFF_SHA1 ""
#define LZ4_SHA1 ""
#define SPDLOG_SHA1 ""
#define XXHASH_SHA1 ""
#define GTEST_SHA1 ""
#define IND_SHA1 ""
#define ROBIN_SHA1 ""
#define TURBOP_SHA1 ""
#define CFR_SHA1 ""
EOF
cat > "$GEN/include/gatb/system/api/config.hpp" <<'EOF'
#define INT128_FOUND            1
#define KSIZE_LIST    32,64
#define KSIZE_STRING "32 64"
#define KSIZE_LIST_TYPE  boost::mpl::int_<32>,boost::mpl::int_<64>
#ifdef GATB_USE_CUSTOM_ALLOCATOR
    #define CUSTOM_MEM_ALLOC  1
#else
    #define CUSTOM_MEM_ALLOC  0
#endif
EOF
cat > "$GEN/include/gatb/system/api/build_info.hpp" <<'EOF'
#define STR_LIBRARY_VERSION     "1.4.1"
#define STR_COMPILATION_DATE    "xxxx-xx-xx"
#define STR_COMPILATION_FLAGS   ""
#define STR_COMPILER            "g++"
#define STR_OPERATING_SYSTEM    "Linux"
EOF
for K in 32 64; do
  for t in TemplateSpecialization1 TemplateSpecialization2; do
    sed "s/\${KSIZE}/$K/g" "$G/src/gatb/template/$t.cpp.in" > "$GEN/src/gatb/template/${t}_$K.cpp"
  done
done
sed -e 's/@KFF_VERSION_MAJOR@/1/' -e 's/@KFF_VERSION_MINOR@/0/' "$T/kff-cpp-api/kff_io.hpp.in" > "$GEN/kff/kff_io.hpp"

# TurboPFor is only reached with --cpr on .hash files (io/hash_file.hpp:100-123,
# out of scope); it builds in-source only, so link abort()ing stubs instead.
cat > "$GEN/turbop_stubs.c" <<'EOF'
#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>
#define STUB(n) size_t n(void* a, size_t b, void* c){(void)a;(void)b;(void)c;fprintf(stderr,"TurboPFor stub " #n " called (--cpr unsupported in oracle/_ref)\n");abort();}
STUB(p4nd1enc64) STUB(p4nzenc8) STUB(p4nzenc16) STUB(p4nzenc32)
STUB(p4nd1dec64) STUB(p4nzdec8) STUB(p4nzdec16) STUB(p4nzdec32)
EOF

INC="-I$REF/include -I$GEN/include -I$GEN/kff -I$T/bcli/include -I$T/indicators/include -I$T/cfrcat/include \
 -I$T/robin-hood-hashing/src/include -I$T/fmt/include -I$T/spdlog/include -I$T/lz4/lib \
 -I$G/src -I$G/thirdparty -I$T/TurboPFor-Integer-Compression/include -I$T/xxHash \
 -I$T/span-lite/include -I$T/bitpacker/include -I$T/googletest/googletest/include"
CXXF="-std=c++17 -O3 -DNDEBUG $MARCH -w -fPIC -DDMAX_C=4294967295 -DWITH_KM_MODULES -DWITH_PLUGIN -DINT128_FOUND"

# ---- object list: "src|obj|lang" ----------------------------------------------------
LIST=$OBJ/list.txt; : > "$LIST"
n=0
for f in $(find "$G/src/gatb" -name '*.cpp' | sort) "$GEN"/src/gatb/template/*.cpp; do
  case "$f" in */kmer/impl/Model.cpp|*/kmer/impl/ConfigurationAlgorithm.cpp|*/kmer/impl/RepartitionAlgorithm.cpp) continue;; esac
  n=$((n+1)); echo "$f|$OBJ/gatb_$n.o|gatb" >> "$LIST"
done
echo "$T/fmt/src/format.cc|$OBJ/fmt_format.o|cxx" >> "$LIST"
echo "$T/fmt/src/os.cc|$OBJ/fmt_os.o|cxx" >> "$LIST"
echo "$T/kff-cpp-api/kff_io.cpp|$OBJ/kff_io.o|cxx" >> "$LIST"
for c in lz4 lz4hc lz4frame xxhash; do echo "$T/lz4/lib/$c.c|$OBJ/lz4_$c.o|c" >> "$LIST"; done
echo "$GEN/turbop_stubs.c|$OBJ/turbop_stubs.o|c" >> "$LIST"
echo "$T/xxHash/xxhash.c|$OBJ/xxhash_main.o|cplain" >> "$LIST"
echo "$REF/src/kmtricks.cpp|$OBJ/km_main.o|cxx" >> "$LIST"
echo "$REF/src/cli.cpp|$OBJ/km_cli.o|cxx" >> "$LIST"
echo "$REF/src/utils.cpp|$OBJ/km_utils.o|cxx" >> "$LIST"
for h in "$HERE"/ref_harness/*.cpp; do
  [ -e "$h" ] && echo "$h|$OBJ/h_$(basename "$h" .cpp).o|cxx" >> "$LIST"
done

compile_one() {
  IFS='|' read -r src obj lang <<< "$1"
  if [ "$obj" -nt "$src" ]; then return 0; fi
  case "$lang" in
    gatb) g++ -std=c++17 -O3 -DNDEBUG $MARCH -w -fPIC -DINT128_FOUND -I"$G/src" -I"$G/thirdparty" -I"$GEN/include" -I"$T/lz4/lib" -c "$src" -o "$obj";;
    cxx)  g++ $CXXF $INC -c "$src" -o "$obj";;
    cplain) gcc -O3 -DNDEBUG $MARCH -w -fPIC -c "$src" -o "$obj";;
    c)    gcc -O3 -DNDEBUG $MARCH -w -fPIC -I"$T/lz4/lib" -DXXH_NAMESPACE=LZ4_ -c "$src" -o "$obj";;
  esac
}
export -f compile_one
export G GEN T INC CXXF MARCH REF
xargs -a "$LIST" -d '\n' -P "$JOBS" -I{} bash -c 'compile_one "$@"' _ {}

LIBOBJS=$(grep -v 'km_main\|km_cli\|km_utils\|/h_' "$LIST" | cut -d'|' -f2 | tr '\n' ' ')
rm -f "$OUT/libkmref$SUFFIX.a"; ar rcs "$OUT/libkmref$SUFFIX.a" $LIBOBJS
g++ -o "$OUT/bin/kmtricks$SUFFIX" "$OBJ/km_main.o" "$OBJ/km_cli.o" "$OBJ/km_utils.o" "$OUT/libkmref$SUFFIX.a" -lz -lpthread -ldl -export-dynamic
if [ -n "$SUFFIX" ]; then echo "oracle/_ref built: bin/kmtricks$SUFFIX ($MARCH)"; exit 0; fi
for h in "$HERE"/ref_harness/*.cpp; do
  [ -e "$h" ] || continue
  b=$(basename "$h" .cpp)
  g++ -o "$OUT/bin/$b" "$OBJ/h_$b.o" "$OUT/libkmref.a" -lz -lpthread -ldl
done
# the reference's example plugins, UNCHANGED (plugins/example/*.cpp): acceptance test of the plugin host ABI
mkdir -p "$OUT/plugins"
for pl in "$REF"/plugins/example/*.cpp; do
  b=$(basename "$pl" .cpp)
  g++ $CXXF $INC -shared "$pl" -o "$OUT/plugins/lib$b.so"
done
echo "oracle/_ref built: $(ls "$OUT/bin" | tr '\n' ' ')"
