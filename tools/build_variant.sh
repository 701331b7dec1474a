#!/bin/bash
# Builds a variant of the library for A/B timing: tools/build_variant.sh NAME [extra nvcc flags...]
# -> build_ab/NAME.so (git-ignored, travels to the GPU box).  Sources are taken from the working tree,
# or from a git revision when J2P_VARIANT_REV is set.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
if [ -n "$J2P_VARIANT_REV" ]; then
  git -C "$root" archive "$J2P_VARIANT_REV" jpeg2png_b200/csrc include | tar -x -C "$tmp"
else
  mkdir -p "$tmp/jpeg2png_b200" && cp -r "$root/jpeg2png_b200/csrc" "$tmp/jpeg2png_b200/" && cp -r "$root/include" "$tmp/"
  rm -f "$tmp"/jpeg2png_b200/csrc/*.o "$tmp"/jpeg2png_b200/csrc/*.so
fi
make -C "$tmp/jpeg2png_b200/csrc" -j8 NVFLAGS_EXTRA="$*" > "$tmp/build.log" 2>&1 || { tail -30 "$tmp/build.log"; exit 1; }
mkdir -p "$root/build_ab"
cp "$tmp/jpeg2png_b200/csrc/libjpeg2png_b200.so" "$root/build_ab/$name.so"
grep -h -A2 "k_gradient_packedILi3ELb1ELi1\|k_project_tmaILb0" "$tmp"/jpeg2png_b200/csrc/*.ptxas.log | grep -E "Used|spill" | head -6
rm -rf "$tmp"
