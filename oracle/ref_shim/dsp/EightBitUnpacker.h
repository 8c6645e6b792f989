// ref_shim/dsp/EightBitUnpacker.h -- TEST INFRASTRUCTURE ONLY: the format unpackers only need HistUnpacker from it.
#include "dsp/HistUnpacker.h"
