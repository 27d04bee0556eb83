// Stand-in for the ITK test driver (CreateTestDriver, /root/reference/test/CMakeLists.txt:6-10): the reference's three test
// programs are compiled UNMODIFIED, from where they lie, against this repo's drop-in headers (include/) and the stand-in ITK
// (oracle/shim); this file only dispatches to them the way `${itk-module}TestDriver <test> <v|fmg|s>` does.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>

using std::abs;  // the tests call abs() unqualified on floats; in an ITK build the floating-point overloads are in scope

int itk2DDiffusionTest_GS(int argc, char* argv[]);
int itk2DDiffusionTest_WJ(int argc, char* argv[]);
int itkVEDTest_GS(int argc, char* argv[]);

int main(int argc, char** argv)
{
  if (argc < 3) { std::fprintf(stderr, "usage: ref_tests_dropin <itk2DDiffusionTest_GS|itk2DDiffusionTest_WJ|itkVEDTest_GS> <v|fmg|s>\n"); return 1; }
  const std::string t = argv[1];
  try {
    if (t == "itk2DDiffusionTest_GS") return itk2DDiffusionTest_GS(argc - 1, argv + 1);
    if (t == "itk2DDiffusionTest_WJ") return itk2DDiffusionTest_WJ(argc - 1, argv + 1);
    if (t == "itkVEDTest_GS") return itkVEDTest_GS(argc - 1, argv + 1);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 2;
  }
  std::fprintf(stderr, "unknown test %s\n", t.c_str());
  return 1;
}
