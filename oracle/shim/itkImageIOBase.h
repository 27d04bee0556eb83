/* stand-in header: nothing of ImageIOBase is used by the reference tests beyond the include (test infrastructure) */
#include "mini_itk.h"
