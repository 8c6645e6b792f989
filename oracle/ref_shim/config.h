/* Stand-in for the autoconf-generated config.h of the reference (not generated here: no autotools).
 * Empty on purpose: HAVE_PSRDADA etc. stay undefined.  TEST INFRASTRUCTURE ONLY. */
