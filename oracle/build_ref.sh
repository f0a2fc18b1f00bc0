#!/bin/bash
# oracle/_ref -- the reference's own hot-path functions, compiled UNMODIFIED from where they lie under /root/reference.
# TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may load the result.
#
# The reference as a whole cannot be built here (OpenCV / Eigen / glog / Sophus / vikit / fast are absent), and its own
# build system is not run.  The three pieces of the hot path that live in the reference's sources need only a few
# members of cv::Mat / cv::KeyPoint / Vector3d, so they are compiled between two stand-in headers of ours:
#
#   ref_stub_prefix.hpp | src/utils.cpp:282-430 | src/initialization.cpp:171-249 | src/camera.cpp:25-41 | ref_stub_suffix.hpp
#
# The reference lines are piped from /root/reference into g++'s stdin: no reference source is copied into this
# repository.  Output: oracle/_ref/libref3dr.so only (git-ignored, travels to the GPU box with the snapshot).
# -O2 -msse2 -ffp-contract=off: the x86-64 baseline the reference builds for (SSE2 half-sampling, no FMA contraction).
set -e
here=$(cd "$(dirname "$0")" && pwd)
ref=${DR3_REFERENCE:-/root/reference}
out=$here/_ref/libref3dr.so
if [ ! -f "$ref/src/utils.cpp" ]; then
    # the GPU box has no /root/reference: the prebuilt library is used as it arrived
    [ -f "$out" ] && { echo "oracle/_ref: reference sources absent, keeping prebuilt $out"; exit 0; }
    echo "oracle/_ref: $ref not present and no prebuilt library" >&2; exit 3
fi
# the line ranges are pinned to the functions they must start / end with: fail loudly if the reference moved
check() { sed -n "$2p" "$ref/$1" | grep -q "$3" || { echo "oracle/_ref: $1:$2 is not '$3'" >&2; exit 4; }; }
check src/utils.cpp 282 'float shi_tomasi_score'
check src/utils.cpp 421 'void create_img_pyramid'
check src/utils.cpp 430 '^}'
check src/initialization.cpp 171 'float InitHelper::CheckFundamental'
check src/initialization.cpp 249 '^}'
check src/camera.cpp 25 'Vector3d Pinhole::cam2world(const double &u, const double &v) const'
check src/camera.cpp 41 '^}'
mkdir -p "$here/_ref"
{
    cat "$here/ref_stub_prefix.hpp"
    echo 'namespace utils {';            sed -n 282,430p "$ref/src/utils.cpp";          echo '}'
    echo 'namespace dr3 { namespace init {'; sed -n 171,249p "$ref/src/initialization.cpp"; echo '} }'
    echo 'namespace dr3 {';              sed -n 25,41p "$ref/src/camera.cpp";           echo '}'
    cat "$here/ref_stub_suffix.hpp"
} | g++ -x c++ -std=c++14 -O2 -msse2 -ffp-contract=off -fno-fast-math -fPIC -shared -o "$out" -
echo "built $out"
