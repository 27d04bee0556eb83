/* stand-in header: see mini_itk_io.h (test infrastructure) */
#include "mini_itk_io.h"
