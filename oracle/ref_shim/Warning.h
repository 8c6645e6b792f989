// ref_shim/Warning.h -- TEST INFRASTRUCTURE ONLY.  Stand-in for PSRCHIVE's Warning (message filter, unused here).
#ifndef REF_SHIM_WARNING_H
#define REF_SHIM_WARNING_H
class Warning {};
#endif
