# oracle/ref.mk -- TEST INFRASTRUCTURE ONLY.
# Compiles, where they lie under /root/reference, the files of the hot path that need nothing outside the
# reference tree except a handful of PSRCHIVE utility headers, for which oracle/ref_shim/ holds stand-ins:
#   libdspsr_refc.so  (plain C)
#     Signal/General/optimize_fft.c   optimal_fft_length            
#     Signal/General/cross_detect.c   cross_detect[_int]            (row a12, Coherence products)
#     Signal/General/stokes_detect.c  stokes_detect[_int]           (row a12, Stokes products)
#     Kernel/Classes/ascii_header.c   ascii_header_get/set          (DADA header keys, Appendix A.8)
#   libdspsr_refcxx.so  (C++, ref_shim/ref_cxx.cpp is the extern "C" door)
#     Kernel/Classes/BitTable.C dsp.C                               (row a1; dsp.C holds the psrdisp_compatible flag)
#     Kernel/Classes/TwoBitTable.C TwoBitLookup.C TwoBitFour.C + dsp/TwoBitFour.h dsp/excision_unpack.h
#       dsp/StepIterator.h                                          (row a6)
#     Signal/General/Dedispersion.C Response.C Shape.C              (rows a7, a8)
# Outputs go to oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).  The rest of the path is C++
# against PSRCHIVE/FFTW class trees (Transformation, TimeSeries, FTransform, Pulsar::Predictor) and cannot be
# built here.  No reference SOURCE is copied into this repository: the compiler reads it in place.
REF ?= /root/reference
CC ?= gcc
CXX ?= g++
OUT = _ref
SRCS = $(REF)/Signal/General/optimize_fft.c $(REF)/Signal/General/cross_detect.c \
       $(REF)/Signal/General/stokes_detect.c $(REF)/Kernel/Classes/ascii_header.c
CXXSRCS = $(REF)/Kernel/Classes/BitTable.C $(REF)/Kernel/Classes/TwoBitTable.C $(REF)/Kernel/Classes/TwoBitLookup.C \
          $(REF)/Kernel/Classes/TwoBitFour.C $(REF)/Kernel/Classes/dsp.C $(REF)/Signal/General/Dedispersion.C $(REF)/Signal/General/Response.C \
          $(REF)/Signal/General/Shape.C
# ref_shim first: its dsp/Observation.h and dsp/ExcisionUnpacker.h stand in for the real ones
INCS = -Iref_shim -I$(REF)/Kernel/Classes -I$(REF)/Signal/General

all: $(OUT)/libdspsr_refc.so $(OUT)/libdspsr_refcxx.so

$(OUT)/libdspsr_refc.so: $(SRCS) ref_shim/config.h
	@mkdir -p $(OUT)
	$(CC) -std=gnu99 -O2 -fPIC -ffp-contract=off -w -shared $(INCS) -o $@ $(SRCS) -lm

# links against the oracle library for the restated JenetAnderson98 numbers only (ref_shim/JenetAnderson98.h)
# the overlap-save loop nests of Filterbank.C / Convolution.C (rows a10, a11), cut out of the files in place into
# _ref/gen/*.inc and compiled inside ref_shim/ref_fbconv.cpp (response multiply = the reference's Response::operate,
# FFT calls = the oracle's restatement of the FFTW conventions).  The greps pin the text to the line numbers.
$(OUT)/gen/filterbank_loop.inc: $(REF)/Signal/General/Filterbank.C $(REF)/Signal/General/Convolution.C
	@mkdir -p $(OUT)/gen
	sed -n '563,660p' $(REF)/Signal/General/Filterbank.C > $(OUT)/gen/filterbank_loop.inc
	sed -n '389,458p' $(REF)/Signal/General/Convolution.C > $(OUT)/gen/convolution_loop.inc
	head -1 $(OUT)/gen/filterbank_loop.inc | grep -q 'for (unsigned input_ichan=0; input_ichan<input->get_nchan(); input_ichan++)'
	grep -q 'backward->bcc1d (freq_res, c_time, freq_dom_ptr);' $(OUT)/gen/filterbank_loop.inc
	grep -q 'data_from = (uint64_t\*)( c_time + nfilt_pos\*2 );' $(OUT)/gen/filterbank_loop.inc
	head -1 $(OUT)/gen/convolution_loop.inc | grep -q 'for (unsigned ichan=0; ichan < nchan; ichan++)'
	grep -q 'memcpy (ptr, complex_time + nfilt_pos\*2, nbytes_step);' $(OUT)/gen/convolution_loop.inc

$(OUT)/libdspsr_refcxx.so: $(CXXSRCS) ref_shim/ref_cxx.cpp ref_shim/ref_fbconv.cpp $(OUT)/gen/filterbank_loop.inc $(wildcard ref_shim/*.h ref_shim/dsp/*.h) $(OUT)/libdspsr_refc.so _build/liboracle.so
	@mkdir -p $(OUT)
	$(CXX) -std=gnu++98 -O2 -fPIC -ffp-contract=off -w -shared $(INCS) -I$(OUT) -o $@ $(CXXSRCS) ref_shim/ref_cxx.cpp ref_shim/ref_fbconv.cpp \
	    -L$(OUT) -ldspsr_refc -L_build -loracle -Wl,-rpath,'$$ORIGIN' -Wl,-rpath,'$$ORIGIN/../_build' -lm

# libdspsr_reffmt.so: the format unpackers (rows a2, a4, a5).  Their own first-level headers are the reference's;
# dsp/EightBitUnpacker.h / dsp/HistUnpacker.h resolve to the ref_shim stand-ins.
FMT = $(REF)/Kernel/Formats
FMTSRCS = $(FMT)/caspsr/CASPSRUnpacker.C $(FMT)/kat/MeerKATUnpacker.C $(FMT)/uwb/UWBUnpacker.C $(REF)/Kernel/Classes/BitTable.C
all: $(OUT)/libdspsr_reffmt.so
$(OUT)/libdspsr_reffmt.so: $(FMTSRCS) ref_shim/ref_formats.cpp $(wildcard ref_shim/*.h ref_shim/dsp/*.h) _build/liboracle.so
	@mkdir -p $(OUT)
	$(CXX) -std=gnu++98 -O2 -fPIC -ffp-contract=off -w -shared -Iref_shim -I$(FMT)/caspsr -I$(FMT)/kat -I$(FMT)/uwb \
	    -I$(REF)/Kernel/Classes -o $@ $(FMTSRCS) ref_shim/ref_formats.cpp \
	    -L_build -loracle -Wl,-rpath,'$$ORIGIN/../_build' -lm -lpthread

# libdspsr_refsigproc.so: the SIGPROC header writer of digifil's last stage, plain C with its state
# in globals (filterbank.h / header.h)
SIGPROC = $(REF)/Kernel/Formats/sigproc
SIGSRCS = $(SIGPROC)/filterbank_header.c $(SIGPROC)/send_stuff.c $(SIGPROC)/strings_equal.c $(SIGPROC)/swap_bytes.c \
          $(SIGPROC)/error_message.c
all: $(OUT)/libdspsr_refsigproc.so
$(OUT)/libdspsr_refsigproc.so: $(SIGSRCS) ref_shim/ref_sigproc.c
	@mkdir -p $(OUT)
	$(CC) -std=gnu99 -O2 -fPIC -w -shared -I$(SIGPROC) -o $@ $(SIGSRCS) ref_shim/ref_sigproc.c -lm

# libdspsr_reffold.so: the loops of dsp::Fold::fold (row a13) compiled from the reference's own text.  Fold.C as a whole
# needs PSRCHIVE's predictor / ephemeris classes; its three self-contained statement blocks do not: sed cuts them, where
# the file lies, into _ref/gen/*.inc (build products, git-ignored) and ref_shim/ref_fold.cpp supplies the variables
# around them.  The grep lines make the build fail if the reference text is not the one the line numbers were read from.
FOLDC = $(REF)/Signal/Pulsar/Fold.C
all: $(OUT)/libdspsr_reffold.so
$(OUT)/gen/fold_binplan.inc: $(FOLDC) $(REF)/Kernel/Classes/WeightedTimeSeries.C
	@mkdir -p $(OUT)/gen
	sed -n '687,716p' $(FOLDC) > $(OUT)/gen/fold_weights.inc
	sed -n '744,787p' $(FOLDC) > $(OUT)/gen/fold_binplan.inc
	sed -n '835,873p' $(FOLDC) > $(OUT)/gen/fold_accum.inc
	grep -q 'iweight = (idat_start + weight_idat) / ndatperweight;' $(OUT)/gen/fold_weights.inc
	head -1 $(OUT)/gen/fold_binplan.inc | grep -q 'for (uint64_t idat=idat_start; idat < idat_end; idat++)'
	grep -q 'phi -= floor(phi);' $(OUT)/gen/fold_binplan.inc
	grep -q 'hits\[ibin\]++;' $(OUT)/gen/fold_binplan.inc
	head -1 $(OUT)/gen/fold_accum.inc | grep -q 'if (in->get_order() == TimeSeries::OrderFPT)'
	grep -q 'phdimp\[idim\] += timep\[idim\];' $(OUT)/gen/fold_accum.inc
	sed -n '584,696p' $(REF)/Kernel/Classes/WeightedTimeSeries.C > $(OUT)/gen/wts_convolve.inc
	sed -n '705,774p' $(REF)/Kernel/Classes/WeightedTimeSeries.C > $(OUT)/gen/wts_scrunch.inc
	head -1 $(OUT)/gen/wts_convolve.inc | grep -q 'if (ndat_per_weight >= nfft)'
	grep -q 'zero_end = uint64_t( ceil((start_idat+nkeep) \* weights_per_dat) );' $(OUT)/gen/wts_convolve.inc
	head -1 $(OUT)/gen/wts_scrunch.inc | grep -q 'uint64_t nweights_tot = get_nweights();'
	grep -q 'weights\[iwt\] /= nscrunch;' $(OUT)/gen/wts_scrunch.inc
$(OUT)/libdspsr_reffold.so: $(OUT)/gen/fold_binplan.inc ref_shim/ref_fold.cpp ref_shim/Error.h
	$(CXX) -std=gnu++98 -O2 -fPIC -ffp-contract=off -w -shared -Iref_shim -I$(OUT) -o $@ ref_shim/ref_fold.cpp -lm

# libdspsr_refbit.so: the generic 8-bit unpacker (row a3): BitUnpacker.C + EightBitUnpacker.C + BitTable.C with their own
# headers; ref_shim/bit comes first so that only dsp/HistUnpacker.h (the PSRCHIVE-dependent base) is a stand-in
BITSRCS = $(REF)/Kernel/Classes/BitUnpacker.C $(REF)/Kernel/Classes/EightBitUnpacker.C $(REF)/Kernel/Classes/BitTable.C
all: $(OUT)/libdspsr_refbit.so
$(OUT)/libdspsr_refbit.so: $(BITSRCS) ref_shim/bit/ref_bitunpack.cpp $(wildcard ref_shim/*.h ref_shim/dsp/*.h ref_shim/bit/dsp/*.h) _build/liboracle.so
	@mkdir -p $(OUT)
	$(CXX) -std=gnu++98 -O2 -fPIC -ffp-contract=off -w -shared -Iref_shim/bit -I$(REF)/Kernel/Classes -Iref_shim \
	    -o $@ $(BITSRCS) ref_shim/bit/ref_bitunpack.cpp -L_build -loracle -Wl,-rpath,'$$ORIGIN/../_build' -lm -lpthread
