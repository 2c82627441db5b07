// TEST STUB -- the few HDF5 C++ class NAMES the reference's headers mention (see hdf5.h here).
#ifndef PB200_TEST_STUB_H5CPP_H
#define PB200_TEST_STUB_H5CPP_H
#include <string>
#include "hdf5.h"
namespace H5 {
class Exception {};
class H5Location {};
class H5File : public H5Location {
  public:
    H5File(int = 0) {}
    hid_t getId() const { return 0; }
};
class Group : public H5Location {};
class DataSet {};
class DataSpace {};
class Attribute {};
}  // namespace H5
#endif
