/* stand-in header: everything lives in mini_itk.h (test infrastructure, see there) */
#include "mini_itk.h"
