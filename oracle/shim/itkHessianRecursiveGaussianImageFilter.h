/* stand-in header: see mini_itk_ved.h (test infrastructure) */
#include "mini_itk_ved.h"
