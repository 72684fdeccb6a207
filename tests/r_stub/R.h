/* stub: see Rinternals.h */
#include "Rinternals.h"
