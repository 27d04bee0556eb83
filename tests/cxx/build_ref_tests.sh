#!/bin/sh
# Builds tests/_build/ref_tests_dropin: the reference's three test programs (UNMODIFIED, compiled from where they lie under
# $REFERENCE/test) against this repo's drop-in headers + the stand-in ITK, linked with libmadgpu.so.  Only possible where
# $REFERENCE exists (the authoring container); the binary travels to the GPU box as a built artefact, like oracle/_ref.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
REFERENCE=${REFERENCE:-/root/reference}
OUT=$ROOT/tests/_build
mkdir -p "$OUT"
for t in itk2DDiffusionTest_GS itk2DDiffusionTest_WJ itkVEDTest_GS; do
  g++ -O1 -std=c++14 -w -include "$HERE/ref_tests_prelude.h" -I"$ROOT/include" -I"$ROOT/oracle/shim" -c -o "$OUT/$t.o" "$REFERENCE/test/$t.cxx"
done
g++ -O1 -std=c++14 -Wall -o "$OUT/ref_tests_dropin" "$HERE/ref_tests_driver.cxx" "$OUT/itk2DDiffusionTest_GS.o" "$OUT/itk2DDiffusionTest_WJ.o" \
    "$OUT/itkVEDTest_GS.o" -L"$ROOT/multigridanisotropicdiffusion_b200" -lmadgpu -lz '-Wl,-rpath,$ORIGIN/../../multigridanisotropicdiffusion_b200'
rm -f "$OUT"/itk2DDiffusionTest_GS.o "$OUT"/itk2DDiffusionTest_WJ.o "$OUT"/itkVEDTest_GS.o
echo "built $OUT/ref_tests_dropin"
