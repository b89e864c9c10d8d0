#!/bin/bash
# oracle/build_ref.sh -- compile the reference spECK (GPUPeople/spECK) in place from
# /root/reference for sm_100 and link it with oracle/ref_wrapper.cu into oracle/_ref/.
# Nothing from the reference is copied into the repository; oracle/_ref/ is git-ignored but
# travels to the GPU box.  Two builds, as the reference's readme (step 5) asks the user to tune
# the shared-memory constants: "stock" (49152/49152) and "tuned" (dynamic = 232448 = B200 opt-in
# maximum), the latter through a generated copy of include/Multiply.h inside oracle/_ref/.
# The vendored 2018 CUB (include/external) is NOT put on the include path: it does not compile
# with the toolkit's Thrust.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${SPECK_REFERENCE_DIR:-/root/reference}
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "build_ref: $REF not present, skipping"; exit 0; }
mkdir -p "$OUT/obj_stock" "$OUT/obj_tuned" "$OUT/include_tuned"
FLAGS="-std=c++17 -O3 -gencode arch=compute_100,code=sm_100 -D_FORCE_INLINES --expt-extended-lambda -use_fast_math --expt-relaxed-constexpr -Xcompiler -fPIC -w"
SRCS="source/GPU/Multiply.cu source/GPU/Compare.cu source/GPU/memory.cpp source/dCSR.cpp source/CSR.cpp source/COO.cpp source/Config.cpp"

build_variant() {  # name, extra include dir (may be empty)
    local name=$1 extra=$2 obj="$OUT/obj_$1"
    local so="$OUT/libspeck_ref_$name.so"
    if [ -f "$so" ] && [ "$so" -nt "$HERE/ref_wrapper.cu" ]; then
        echo "build_ref: $so up to date"; return
    fi
    local inc="-I$REF/include -I$REF/externals"
    [ -n "$extra" ] && inc="-I$extra $inc"
    local pids=""
    for s in $SRCS; do   # the reference's objects are rebuilt only when missing (its sources are read-only)
        o="$obj/$(basename ${s%.*}).o"
        if [ ! -f "$o" ]; then
            nvcc $FLAGS $inc -x cu -c "$REF/$s" -o "$o" &
            pids="$pids $!"
        fi
    done
    nvcc $FLAGS $inc -c "$HERE/ref_wrapper.cu" -o "$obj/ref_wrapper.o" &
    pids="$pids $!"
    for p in $pids; do wait $p; done
    nvcc -gencode arch=compute_100,code=sm_100 -shared -o "$so" "$obj"/*.o
    echo "build_ref: built $so"
}

# tuned header: only the dynamic shared-memory constant changes (readme.md step 5)
sed 's/spECK_DYNAMIC_MEM_PER_BLOCK{49152}/spECK_DYNAMIC_MEM_PER_BLOCK{232448}/' "$REF/include/Multiply.h" > "$OUT/include_tuned/Multiply.h"
grep -q 232448 "$OUT/include_tuned/Multiply.h" || { echo "build_ref: could not patch the tuned header"; exit 1; }

build_variant stock "" &
P1=$!
build_variant tuned "$OUT/include_tuned" &
P2=$!
wait $P1; wait $P2
ls -la "$OUT"/*.so
