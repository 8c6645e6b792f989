// ref_shim/OwnStream.h -- TEST INFRASTRUCTURE ONLY.  Classes deriving from OwnStream own a `cerr` stream.
#ifndef REF_SHIM_OWNSTREAM_H
#define REF_SHIM_OWNSTREAM_H
#include <iostream>
#include "Reference.h"
#include "environ.h"
class OwnStream : public Reference::Able {
 public:
  OwnStream() : cerr(std::cerr.rdbuf()) {}
  OwnStream(const OwnStream&) : Reference::Able(), cerr(std::cerr.rdbuf()) {}
  const OwnStream& operator=(const OwnStream&) { return *this; }
  virtual ~OwnStream() {}
  virtual void set_cerr(std::ostream& os) const { cerr.rdbuf(os.rdbuf()); }
 protected:
  mutable std::ostream cerr;
};
#endif
