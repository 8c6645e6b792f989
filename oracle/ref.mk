# oracle/ref.mk -- TEST INFRASTRUCTURE ONLY.
# Compiles, where they lie under /root/reference, the few files of the hot path that are plain C
# and need nothing outside the reference tree:
#   Signal/General/optimize_fft.c   optimal_fft_length            
#   Signal/General/cross_detect.c   cross_detect[_int]            (row a12, Coherence products)
#   Signal/General/stokes_detect.c  stokes_detect[_int]           (row a12, Stokes products)
#   Kernel/Classes/ascii_header.c   ascii_header_get/set          (DADA header keys, Appendix A.8)
# into oracle/_ref/libdspsr_refc.so (git-ignored; travels to the GPU box with the snapshot).
# Everything else on the path is C++ against PSRCHIVE/FFTW and cannot be built here.
# No reference SOURCE is copied into this repository: the compiler reads it in place.
REF ?= /root/reference
CC ?= gcc
OUT = _ref
SRCS = $(REF)/Signal/General/optimize_fft.c $(REF)/Signal/General/cross_detect.c \
       $(REF)/Signal/General/stokes_detect.c $(REF)/Kernel/Classes/ascii_header.c

all: $(OUT)/libdspsr_refc.so

$(OUT)/libdspsr_refc.so: $(SRCS) ref_shim/config.h
	@mkdir -p $(OUT)
	$(CC) -std=gnu99 -O2 -fPIC -ffp-contract=off -w -shared -Iref_shim -I$(REF)/Kernel/Classes -I$(REF)/Signal/General \
	    -o $@ $(SRCS) -lm
