// ref_shim/bit/dsp/HistUnpacker.h -- TEST INFRASTRUCTURE ONLY.  For libdspsr_refbit.so the include path is
// ref_shim/bit, then the reference's Kernel/Classes, then ref_shim: "dsp/HistUnpacker.h" is the stand-in while
// "dsp/BitUnpacker.h" and "dsp/EightBitUnpacker.h" are the reference's own headers.
#include "../../dsp/HistUnpacker.h"
