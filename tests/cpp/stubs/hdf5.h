// TEST STUB -- just enough of the HDF5 C API's TYPES for the reference's optimisation headers
// (include/parpeoptimization/optimizationProblem.h:11, include/parpecommon/hdf5Misc.h:5-7) to be
// parsed on a box without an HDF5 installation (SURVEY.md hard part H6).  Nothing here is called:
// tests/cpp/test_minibatch.cpp drives the reference's unmodified MinibatchOptimizer without a
// reporter / result writer.
#ifndef PB200_TEST_STUB_HDF5_H
#define PB200_TEST_STUB_HDF5_H
#include <stddef.h>
#include <stdint.h>
typedef int64_t hid_t;
typedef int herr_t;
typedef unsigned long long hsize_t;
typedef int htri_t;
typedef struct H5E_error_t { int dummy; } H5E_error_t;
typedef struct H5E_error2_t { int dummy; } H5E_error2_t;
typedef herr_t (*H5E_auto1_t)(void*);
typedef herr_t (*H5E_auto2_t)(hid_t, void*);
inline herr_t H5Eget_auto1(H5E_auto1_t*, void**) { return 0; }
inline herr_t H5Eset_auto1(H5E_auto1_t, void*) { return 0; }
#define H5E_DEFAULT 0
#endif
