// ref_shim/environ.h -- TEST INFRASTRUCTURE ONLY.  Fixed-width integers and the printf macros of PSRCHIVE's environ.h.
#ifndef REF_SHIM_ENVIRON_H
#define REF_SHIM_ENVIRON_H
#include <stdint.h>
#include <inttypes.h>
#define I64 "%" PRIi64
#define UI64 "%" PRIu64
#endif
