// ref_shim/FTransformBench.h -- TEST INFRASTRUCTURE ONLY (the benchmark-driven FFT length policy is not exercised).
#ifndef REF_SHIM_FTRANSFORMBENCH_H
#define REF_SHIM_FTRANSFORMBENCH_H
#include "Reference.h"
namespace FTransform { class Bench : public Reference::Able {}; }
#endif
