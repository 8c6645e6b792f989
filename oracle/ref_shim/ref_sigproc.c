/* TEST INFRASTRUCTURE: the one global of the reference's sigproc C files that is defined in a C++ file
 * (Kernel/Formats/sigproc/SigProcObservation.C) which needs PSRCHIVE and cannot be compiled here. */
char sigproc_verbose = 0;
