// TEST STUB (see hdf5.h in this directory)
#include "hdf5.h"
