/* Stand-in for the autoconf-generated config.h of the reference build
 * (test infrastructure: used only by oracle/Makefile.ref to compile the
 * unmodified reference sources under /root/reference into oracle/_ref/). */
#ifndef AWM_REF_SHIM_CONFIG_H
#define AWM_REF_SHIM_CONFIG_H
#define VERSION "0.6.5-refshim"
#define PACKAGE_VERSION VERSION
#endif
