#!/bin/bash
# integration/build.sh -- build the reference's host pipeline around the drop-in boundary.
#
# Compiles the reference's own host sources where they lie under $REF_ROOT (CMakeLists.txt:71-93 minus
# src/gasal2_ssw.cpp; nothing is copied into this repo) and links them three ways into integration/_build/
# (git-ignored, travels to the GPU box with the snapshot):
#
#   rabbitsalign_b200     integration/gasal2_ssw.cpp + librsa_ext.so     the product drop-in (needs a GPU)
#   rabbitsalign_gasalref oracle/_ref/libgasal_ref512.so behind solve_ssw_on_gpu: the reference's own GASAL2
#                         kernels compiled for the host = golden-SAM generator (CPU only, test infrastructure)
#   rabbitsalign_b200_alninfo  the product drop-in plus the optional caller-loop edit (patch_caller.py): gasal_fail +
#                         Aligner::align_gpu come from the device (finish_kernel); needs a GPU
#   rabbitsalign_gasalgpu the reference as shipped: its own solve_ssw_on_gpu + GASAL2 for sm_100a (comparator, needs a GPU)
#   rabbitsalign_cpussw   solve_ssw_on_gpu returns failed records -> every extension takes the reference's
#                         CPU SSW path (Aligner::align): the end-to-end CPU baseline (test/bench infrastructure)
#
# The reference's src/gasal2_ssw.h is pre-empted by force-including integration/gasal2_ssw.h (same include
# guard), so src/pc.cpp, src/aligner.hpp and ext/ssw/ssw_cpp.h compile unedited.  Three build-time stand-ins
# are generated under _build/gen/: zstr.hpp (the reference fetches zstr from GitHub at configure time,
# CMakeLists.txt:32-37; plain std::ifstream is enough for uncompressed FASTA), version.hpp and buildconfig.hpp
# (configured from src/*.in by CMake).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF_ROOT="${REF_ROOT:-/root/reference}"
OUT="$HERE/_build"
CXX="${CXX:-/usr/bin/g++}"
CC="${CC:-/usr/bin/gcc}"
[ -d "$REF_ROOT/src" ] || { echo "integration/build.sh: $REF_ROOT absent - keeping prebuilt _build (if any)"; exit 0; }
mkdir -p "$OUT/gen" "$OUT/obj"

cat > "$OUT/gen/zstr.hpp" <<'EOS'
#pragma once
#include <fstream>
namespace zstr { using ifstream = std::ifstream; }
EOS
sed 's/@PROJECT_VERSION@/0.11.0-b200/' "$REF_ROOT/src/version.hpp.in" > "$OUT/gen/version.hpp"
sed 's/@CMAKE_BUILD_TYPE@/Release/' "$REF_ROOT/src/buildconfig.hpp.in" > "$OUT/gen/buildconfig.hpp"

FLAGS="-O3 -march=x86-64-v3 -std=c++17 -w -pthread -DNDEBUG -I$OUT/gen -I$REF_ROOT/src -I$REF_ROOT/ext -I$ROOT/include -I$HERE -include $HERE/gasal2_ssw.h"
SRCS="refs fastq cmdline index indexparameters sam paf pc aln cigar aligner nam randstrobes readlen version io main"
pids=()
for s in $SRCS; do
  ( [ "$OUT/obj/$s.o" -nt "$REF_ROOT/src/$s.cpp" ] && [ "$OUT/obj/$s.o" -nt "$HERE/gasal2_ssw.h" ] || $CXX $FLAGS -c "$REF_ROOT/src/$s.cpp" -o "$OUT/obj/$s.o" ) &
  pids+=($!)
done
( $CXX $FLAGS -c "$REF_ROOT/ext/ssw/ssw_cpp.cpp" -o "$OUT/obj/ssw_cpp.o" ) & pids+=($!)
( $CC -O3 -march=x86-64-v3 -w -c "$REF_ROOT/ext/ssw/ssw.c" -o "$OUT/obj/ssw.o" ) & pids+=($!)
( $CC -O3 -march=x86-64-v3 -w -c "$REF_ROOT/ext/xxhash.c" -o "$OUT/obj/xxhash.o" ) & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done

OBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do echo "$OUT/obj/$s.o"; done)
# 1) the product drop-in
$CXX $FLAGS -c "$HERE/gasal2_ssw.cpp" -o "$OUT/obj/veneer.o"
$CXX -o "$OUT/rabbitsalign_b200" $OBJS "$OUT/obj/veneer.o" -L"$ROOT/rabbitsalign_b200" -lrsa_ext \
     -Wl,-rpath,'$ORIGIN/../../rabbitsalign_b200' -lz -lpthread
# 2) golden generator: the reference's kernels compiled for the host
$CXX $FLAGS -c "$HERE/solve_gasalref.cpp" -o "$OUT/obj/solve_gasalref.o"
$CXX -o "$OUT/rabbitsalign_gasalref" $OBJS "$OUT/obj/solve_gasalref.o" "$ROOT/oracle/_ref/libgasal_ref512.so" \
     -Wl,-rpath,'$ORIGIN/../../oracle/_ref' -lz -lpthread
# 2b) the reference AS SHIPPED, on this GPU: its own solve_ssw_on_gpu + GASAL2 compiled for sm_100a
#     (oracle/_ref/libgasal_gpu.so, oracle/Makefile): the end-to-end GPU comparator (needs a GPU)
if [ -f "$ROOT/oracle/_ref/libgasal_gpu.so" ]; then
  $CXX -o "$OUT/rabbitsalign_gasalgpu" $OBJS "$ROOT/oracle/_ref/libgasal_gpu.so" \
       -Wl,-rpath,'$ORIGIN/../../oracle/_ref' -lz -lpthread
fi
# 3) CPU-SSW path
$CXX $FLAGS -c "$HERE/solve_cpussw.cpp" -o "$OUT/obj/solve_cpussw.o"
$CXX -o "$OUT/rabbitsalign_cpussw" $OBJS "$OUT/obj/solve_cpussw.o" -lz -lpthread

# ---- the reference's SHIPPED configuration: RabbitFX chunk reader + OPT_NUMA_CLOSE (build.sh:49
#      -DUSE_RABBITFX=ON -DCLOSE_NUMA_OPT=ON).  RabbitFX/io/*.cpp are compiled directly (its own CMake needs
#      FindOpenMP), with `-include cstdint` for GCC 13; src/pc.cpp and src/main.cpp are rebuilt with the two macros.
FXFLAGS="$FLAGS -DRABBIT_FX -DOPT_NUMA_CLOSE -DVERB -include cstdint -I$REF_ROOT/RabbitFX/io"
mkdir -p "$OUT/obj_fx"
pids=()
for s in FastxStream Formater verbos_Formater; do
  ( [ "$OUT/obj_fx/$s.o" -nt "$REF_ROOT/RabbitFX/io/$s.cpp" ] || $CXX -O3 -std=c++17 -w -DVERB -include cstdint -I"$REF_ROOT/RabbitFX/io" -c "$REF_ROOT/RabbitFX/io/$s.cpp" -o "$OUT/obj_fx/$s.o" ) &
  pids+=($!)
done
for s in pc main; do
  ( [ "$OUT/obj_fx/$s.o" -nt "$REF_ROOT/src/$s.cpp" ] && [ "$OUT/obj_fx/$s.o" -nt "$HERE/gasal2_ssw.h" ] || $CXX $FXFLAGS -c "$REF_ROOT/src/$s.cpp" -o "$OUT/obj_fx/$s.o" ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
FXOBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do if [ "$s" = pc ] || [ "$s" = main ]; then echo "$OUT/obj_fx/$s.o"; else echo "$OUT/obj/$s.o"; fi; done)
# like the reference (RabbitFX/io/CMakeLists.txt:19-20): a static archive, members pulled on demand
rm -f "$OUT/obj_fx/librabbitfx.a"
ar rcs "$OUT/obj_fx/librabbitfx.a" "$OUT/obj_fx/FastxStream.o" "$OUT/obj_fx/Formater.o" "$OUT/obj_fx/verbos_Formater.o"
FXIO="$OUT/obj_fx/librabbitfx.a"
$CXX -o "$OUT/rabbitsalign_fx_b200" $FXOBJS $FXIO "$OUT/obj/veneer.o" -L"$ROOT/rabbitsalign_b200" -lrsa_ext \
     -Wl,-rpath,'$ORIGIN/../../rabbitsalign_b200' -lz -lpthread
$CXX -o "$OUT/rabbitsalign_fx_gasalref" $FXOBJS $FXIO "$OUT/obj/solve_gasalref.o" "$ROOT/oracle/_ref/libgasal_ref512.so" \
     -Wl,-rpath,'$ORIGIN/../../oracle/_ref' -lz -lpthread
$CXX -o "$OUT/rabbitsalign_fx_cpussw" $FXOBJS $FXIO "$OUT/obj/solve_cpussw.o" -lz -lpthread
if [ -f "$ROOT/oracle/_ref/libgasal_gpu.so" ]; then
  $CXX -o "$OUT/rabbitsalign_fx_gasalgpu" $FXOBJS $FXIO "$ROOT/oracle/_ref/libgasal_gpu.so" \
       -Wl,-rpath,'$ORIGIN/../../oracle/_ref' -lz -lpthread
fi

# ---- optional build: AlignmentInfo straight from the device.  The caller loops of src/pc.cpp get
#      the two-call substitution of integration/patch_caller.py on a copy under _build/alninfo/; every unit that
#      sees gasal_tmp_res is rebuilt with -DRSA_EXT_ALNINFO (the record carries the device's rsa_ext_alninfo_t).
mkdir -p "$OUT/alninfo" "$OUT/obj_aln"
python3 "$HERE/patch_caller.py" "$REF_ROOT/src/pc.cpp" "$OUT/alninfo/pc.cpp"
ALNFLAGS="$FLAGS -DRSA_EXT_ALNINFO"
pids=()
( $CXX $ALNFLAGS -c "$OUT/alninfo/pc.cpp" -o "$OUT/obj_aln/pc.o" ) & pids+=($!)
( $CXX $ALNFLAGS -DRABBIT_FX -DOPT_NUMA_CLOSE -DVERB -include cstdint -I"$REF_ROOT/RabbitFX/io" -c "$OUT/alninfo/pc.cpp" -o "$OUT/obj_aln/pc_fx.o" ) & pids+=($!)
( $CXX $ALNFLAGS -c "$REF_ROOT/src/aligner.cpp" -o "$OUT/obj_aln/aligner.o" ) & pids+=($!)
( $CXX $ALNFLAGS -c "$REF_ROOT/ext/ssw/ssw_cpp.cpp" -o "$OUT/obj_aln/ssw_cpp.o" ) & pids+=($!)
( $CXX $ALNFLAGS -c "$HERE/gasal2_ssw.cpp" -o "$OUT/obj_aln/veneer.o" ) & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
ALNOBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do case $s in pc|aligner|ssw_cpp) echo "$OUT/obj_aln/$s.o";; *) echo "$OUT/obj/$s.o";; esac; done)
$CXX -o "$OUT/rabbitsalign_b200_alninfo" $ALNOBJS "$OUT/obj_aln/veneer.o" -L"$ROOT/rabbitsalign_b200" -lrsa_ext \
     -Wl,-rpath,'$ORIGIN/../../rabbitsalign_b200' -lz -lpthread
ALNFXOBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do case $s in pc) echo "$OUT/obj_aln/pc_fx.o";; main) echo "$OUT/obj_fx/main.o";; aligner|ssw_cpp) echo "$OUT/obj_aln/$s.o";; *) echo "$OUT/obj/$s.o";; esac; done)
$CXX -o "$OUT/rabbitsalign_fx_b200_alninfo" $ALNFXOBJS $FXIO "$OUT/obj_aln/veneer.o" -L"$ROOT/rabbitsalign_b200" -lrsa_ext \
     -Wl,-rpath,'$ORIGIN/../../rabbitsalign_b200' -lz -lpthread
# ---- zero-edit whole-chunk build (VERDICT r1 task 1a): the same unmodified sources, -DSTREAM_BATCH_SIZE=1048576, so a
#      chunk's todo list goes down in one solve_ssw_on_gpu call instead of 512-pair slices (the engine takes any n).
mkdir -p "$OUT/obj_big"
BIGFLAGS="$FLAGS -DSTREAM_BATCH_SIZE=1048576"
pids=()
( $CXX $BIGFLAGS -c "$REF_ROOT/src/pc.cpp" -o "$OUT/obj_big/pc.o" ) & pids+=($!)
( $CXX $BIGFLAGS -DRABBIT_FX -DOPT_NUMA_CLOSE -DVERB -include cstdint -I"$REF_ROOT/RabbitFX/io" -c "$REF_ROOT/src/pc.cpp" -o "$OUT/obj_big/pc_fx.o" ) & pids+=($!)
( $CXX $BIGFLAGS -c "$HERE/gasal2_ssw.cpp" -o "$OUT/obj_big/veneer.o" ) & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
BIGOBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do case $s in pc) echo "$OUT/obj_big/pc.o";; *) echo "$OUT/obj/$s.o";; esac; done)
$CXX -o "$OUT/rabbitsalign_b200_big" $BIGOBJS "$OUT/obj_big/veneer.o" -L"$ROOT/rabbitsalign_b200" -lrsa_ext \
     -Wl,-rpath,'$ORIGIN/../../rabbitsalign_b200' -lz -lpthread
BIGFXOBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do case $s in pc) echo "$OUT/obj_big/pc_fx.o";; main) echo "$OUT/obj_fx/main.o";; *) echo "$OUT/obj/$s.o";; esac; done)
$CXX -o "$OUT/rabbitsalign_fx_b200_big" $BIGFXOBJS $FXIO "$OUT/obj_big/veneer.o" -L"$ROOT/rabbitsalign_b200" -lrsa_ext \
     -Wl,-rpath,'$ORIGIN/../../rabbitsalign_b200' -lz -lpthread

# ---- window build (caller half): patch_caller.py --windows on a copy of src/pc.cpp:
#      windows travel as (contig, start, length), one call per chunk, genome resident in HBM, AlignmentInfo from the device.
mkdir -p "$OUT/windows" "$OUT/obj_win"
python3 "$HERE/patch_caller.py" --windows "$REF_ROOT/src/pc.cpp" "$OUT/windows/pc.cpp"
WINFLAGS="$FLAGS -DRSA_EXT_WINDOWS"
pids=()
( $CXX $WINFLAGS -c "$OUT/windows/pc.cpp" -o "$OUT/obj_win/pc.o" ) & pids+=($!)
( $CXX $WINFLAGS -DRABBIT_FX -DOPT_NUMA_CLOSE -DVERB -include cstdint -I"$REF_ROOT/RabbitFX/io" -c "$OUT/windows/pc.cpp" -o "$OUT/obj_win/pc_fx.o" ) & pids+=($!)
( $CXX $WINFLAGS -c "$HERE/gasal2_ssw.cpp" -o "$OUT/obj_win/veneer.o" ) & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
# (aligner.o / ssw_cpp.o of the alninfo build: the record layout is the same)
WINOBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do case $s in pc) echo "$OUT/obj_win/pc.o";; aligner|ssw_cpp) echo "$OUT/obj_aln/$s.o";; *) echo "$OUT/obj/$s.o";; esac; done)
$CXX -o "$OUT/rabbitsalign_b200_win" $WINOBJS "$OUT/obj_win/veneer.o" -L"$ROOT/rabbitsalign_b200" -lrsa_ext \
     -Wl,-rpath,'$ORIGIN/../../rabbitsalign_b200' -lz -lpthread
WINFXOBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do case $s in pc) echo "$OUT/obj_win/pc_fx.o";; main) echo "$OUT/obj_fx/main.o";; aligner|ssw_cpp) echo "$OUT/obj_aln/$s.o";; *) echo "$OUT/obj/$s.o";; esac; done)
$CXX -o "$OUT/rabbitsalign_fx_b200_win" $WINFXOBJS $FXIO "$OUT/obj_win/veneer.o" -L"$ROOT/rabbitsalign_b200" -lrsa_ext \
     -Wl,-rpath,'$ORIGIN/../../rabbitsalign_b200' -lz -lpthread
# ---- GPU seeding build: the window build + patch_seed.py: a chunk's reads are seeded
#      in one rsa_seed_find_nams call (randstrobes, index lookup, NAM merge, rescue on the GPU); the per-read loop of the
#      reference consumes the precomputed NAM lists.  Index replicated per GPU, shared by the workers.
mkdir -p "$OUT/seed" "$OUT/obj_seed"
python3 "$HERE/patch_seed.py" "$REF_ROOT/src/aln.cpp" "$OUT/seed/aln.cpp" "$OUT/windows/pc.cpp" "$OUT/seed/pc.cpp"
pids=()
( $CXX $WINFLAGS -c "$OUT/seed/aln.cpp" -o "$OUT/obj_seed/aln.o" ) & pids+=($!)
( $CXX $WINFLAGS -c "$OUT/seed/pc.cpp" -o "$OUT/obj_seed/pc.o" ) & pids+=($!)
( $CXX $WINFLAGS -DRABBIT_FX -DOPT_NUMA_CLOSE -DVERB -include cstdint -I"$REF_ROOT/RabbitFX/io" -c "$OUT/seed/pc.cpp" -o "$OUT/obj_seed/pc_fx.o" ) & pids+=($!)
( $CXX $WINFLAGS -c "$HERE/seed_glue.cpp" -o "$OUT/obj_seed/seed_glue.o" ) & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
SEEDOBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do case $s in pc) echo "$OUT/obj_seed/pc.o";; aln) echo "$OUT/obj_seed/aln.o";; aligner|ssw_cpp) echo "$OUT/obj_aln/$s.o";; *) echo "$OUT/obj/$s.o";; esac; done)
$CXX -o "$OUT/rabbitsalign_b200_gpuseed" $SEEDOBJS "$OUT/obj_seed/seed_glue.o" "$OUT/obj_win/veneer.o" -L"$ROOT/rabbitsalign_b200" -lrsa_ext \
     -Wl,-rpath,'$ORIGIN/../../rabbitsalign_b200' -lz -lpthread
SEEDFXOBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do case $s in pc) echo "$OUT/obj_seed/pc_fx.o";; aln) echo "$OUT/obj_seed/aln.o";; main) echo "$OUT/obj_fx/main.o";; aligner|ssw_cpp) echo "$OUT/obj_aln/$s.o";; *) echo "$OUT/obj/$s.o";; esac; done)
$CXX -o "$OUT/rabbitsalign_fx_b200_gpuseed" $SEEDFXOBJS $FXIO "$OUT/obj_seed/seed_glue.o" "$OUT/obj_win/veneer.o" -L"$ROOT/rabbitsalign_b200" -lrsa_ext \
     -Wl,-rpath,'$ORIGIN/../../rabbitsalign_b200' -lz -lpthread
# ---- Hamming-on-device build (caller half): the GPU seeding build + patch_hamming.py: the
#      Hamming shortcut of extend_seed_part is decided for a whole chunk in one rsa_ext_hamming_ref_windows call (windows read
#      from the genome resident in HBM); candidates that fail the 5 % test join the chunk's Smith-Waterman batch as before.
mkdir -p "$OUT/ham" "$OUT/obj_ham"
python3 "$HERE/patch_hamming.py" "$OUT/seed/aln.cpp" "$OUT/ham/aln.cpp" "$OUT/seed/pc.cpp" "$OUT/ham/pc.cpp"
pids=()
( $CXX $WINFLAGS -c "$OUT/ham/aln.cpp" -o "$OUT/obj_ham/aln.o" ) & pids+=($!)
( $CXX $WINFLAGS -c "$OUT/ham/pc.cpp" -o "$OUT/obj_ham/pc.o" ) & pids+=($!)
( $CXX $WINFLAGS -DRABBIT_FX -DOPT_NUMA_CLOSE -DVERB -include cstdint -I"$REF_ROOT/RabbitFX/io" -c "$OUT/ham/pc.cpp" -o "$OUT/obj_ham/pc_fx.o" ) & pids+=($!)
( $CXX $WINFLAGS -c "$HERE/hamming_glue.cpp" -o "$OUT/obj_ham/hamming_glue.o" ) & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
HAMOBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do case $s in pc) echo "$OUT/obj_ham/pc.o";; aln) echo "$OUT/obj_ham/aln.o";; aligner|ssw_cpp) echo "$OUT/obj_aln/$s.o";; *) echo "$OUT/obj/$s.o";; esac; done)
$CXX -o "$OUT/rabbitsalign_b200_gpuham" $HAMOBJS "$OUT/obj_ham/hamming_glue.o" "$OUT/obj_seed/seed_glue.o" "$OUT/obj_win/veneer.o" -L"$ROOT/rabbitsalign_b200" -lrsa_ext \
     -Wl,-rpath,'$ORIGIN/../../rabbitsalign_b200' -lz -lpthread
HAMFXOBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do case $s in pc) echo "$OUT/obj_ham/pc_fx.o";; aln) echo "$OUT/obj_ham/aln.o";; main) echo "$OUT/obj_fx/main.o";; aligner|ssw_cpp) echo "$OUT/obj_aln/$s.o";; *) echo "$OUT/obj/$s.o";; esac; done)
$CXX -o "$OUT/rabbitsalign_fx_b200_gpuham" $HAMFXOBJS $FXIO "$OUT/obj_ham/hamming_glue.o" "$OUT/obj_seed/seed_glue.o" "$OUT/obj_win/veneer.o" -L"$ROOT/rabbitsalign_b200" -lrsa_ext \
     -Wl,-rpath,'$ORIGIN/../../rabbitsalign_b200' -lz -lpthread
# ---- SAM-from-the-device build (caller half): the Hamming build + patch_sam.py: the three
#      text-appending members of class Sam hand their arguments to a per-chunk collector, one rsa_sam_format call per chunk
#      writes the text.  With this build every step between FASTQ parsing and SAM bytes that touches bases runs on the GPU.
mkdir -p "$OUT/sam" "$OUT/obj_sam"
python3 "$HERE/patch_sam.py" "$REF_ROOT/src/sam.cpp" "$OUT/sam/sam.cpp" "$OUT/ham/pc.cpp" "$OUT/sam/pc.cpp"
pids=()
( $CXX $WINFLAGS -c "$OUT/sam/sam.cpp" -o "$OUT/obj_sam/sam.o" ) & pids+=($!)
( $CXX $WINFLAGS -c "$OUT/sam/pc.cpp" -o "$OUT/obj_sam/pc.o" ) & pids+=($!)
( $CXX $WINFLAGS -DRABBIT_FX -DOPT_NUMA_CLOSE -DVERB -include cstdint -I"$REF_ROOT/RabbitFX/io" -c "$OUT/sam/pc.cpp" -o "$OUT/obj_sam/pc_fx.o" ) & pids+=($!)
( $CXX $WINFLAGS -c "$HERE/sam_glue.cpp" -o "$OUT/obj_sam/sam_glue.o" ) & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
SAMOBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do case $s in pc) echo "$OUT/obj_sam/pc.o";; sam) echo "$OUT/obj_sam/sam.o";; aln) echo "$OUT/obj_ham/aln.o";; aligner|ssw_cpp) echo "$OUT/obj_aln/$s.o";; *) echo "$OUT/obj/$s.o";; esac; done)
$CXX -o "$OUT/rabbitsalign_b200_gpusam" $SAMOBJS "$OUT/obj_sam/sam_glue.o" "$OUT/obj_ham/hamming_glue.o" "$OUT/obj_seed/seed_glue.o" "$OUT/obj_win/veneer.o" -L"$ROOT/rabbitsalign_b200" -lrsa_ext \
     -Wl,-rpath,'$ORIGIN/../../rabbitsalign_b200' -lz -lpthread
SAMFXOBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do case $s in pc) echo "$OUT/obj_sam/pc_fx.o";; sam) echo "$OUT/obj_sam/sam.o";; aln) echo "$OUT/obj_ham/aln.o";; main) echo "$OUT/obj_fx/main.o";; aligner|ssw_cpp) echo "$OUT/obj_aln/$s.o";; *) echo "$OUT/obj/$s.o";; esac; done)
$CXX -o "$OUT/rabbitsalign_fx_b200_gpusam" $SAMFXOBJS $FXIO "$OUT/obj_sam/sam_glue.o" "$OUT/obj_ham/hamming_glue.o" "$OUT/obj_seed/seed_glue.o" "$OUT/obj_win/veneer.o" -L"$ROOT/rabbitsalign_b200" -lrsa_ext \
     -Wl,-rpath,'$ORIGIN/../../rabbitsalign_b200' -lz -lpthread
# ---- profiling builds: the reference's own per-phase timers (time1 seeding loop, time2_1..4 extension phases, time3_1..2
#      SAM + output; their summary fprintf is commented out in src/pc.cpp:806-809 and siblings) switched back on
mkdir -p "$OUT/timed"
sed -E '/^ *\/\/fprintf\($/,/^ *\/\/\);$/ s,^( *)//,\1,' "$REF_ROOT/src/pc.cpp" > "$OUT/timed/pc_ref.cpp"
sed -E '/^ *\/\/fprintf\($/,/^ *\/\/\);$/ s,^( *)//,\1,' "$OUT/seed/pc.cpp" > "$OUT/timed/pc_seed.cpp"
pids=()
( $CXX $FLAGS -c "$OUT/timed/pc_ref.cpp" -o "$OUT/timed/pc_ref.o" ) & pids+=($!)
( $CXX $WINFLAGS -c "$OUT/timed/pc_seed.cpp" -o "$OUT/timed/pc_seed.o" ) & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
if [ -f "$ROOT/oracle/_ref/libgasal_gpu.so" ]; then
  TOBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do case $s in pc) echo "$OUT/timed/pc_ref.o";; *) echo "$OUT/obj/$s.o";; esac; done)
  $CXX -o "$OUT/rabbitsalign_gasalgpu_timed" $TOBJS "$ROOT/oracle/_ref/libgasal_gpu.so" -Wl,-rpath,'$ORIGIN/../../oracle/_ref' -lz -lpthread
fi
TSOBJS=$(for s in $SRCS ssw_cpp ssw xxhash; do case $s in pc) echo "$OUT/timed/pc_seed.o";; aln) echo "$OUT/obj_seed/aln.o";; aligner|ssw_cpp) echo "$OUT/obj_aln/$s.o";; *) echo "$OUT/obj/$s.o";; esac; done)
$CXX -o "$OUT/rabbitsalign_b200_gpuseed_timed" $TSOBJS "$OUT/obj_seed/seed_glue.o" "$OUT/obj_win/veneer.o" -L"$ROOT/rabbitsalign_b200" -lrsa_ext \
     -Wl,-rpath,'$ORIGIN/../../rabbitsalign_b200' -lz -lpthread
ls -la "$OUT"/rabbitsalign_*
