// Force-included in front of the reference's unmodified test programs (tests/test_ref_tests_dropin.py): they use std::strcmp and
// unqualified abs() on floats without including <cstring> / relying on ITK's headers to have the floating-point overloads in
// scope.  A real ITK build provides both transitively; the stand-in ITK says so here.
#include <cmath>
#include <cstdlib>
#include <cstring>
using std::abs;
