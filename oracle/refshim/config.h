/* Hand-written stand-in for the autotools-generated config.h of the reference
 * (values from /root/reference/configure.ac:5,16-19).  Used ONLY when compiling
 * the reference's own hot-path translation units into oracle/_ref/ as the parity
 * oracle; never part of the product. */
#ifndef CRASS_ORACLE_REF_CONFIG_H
#define CRASS_ORACLE_REF_CONFIG_H
#define PACKAGE_NAME "crass"
#define PACKAGE_VERSION "1.0.1"
#define PACKAGE_FULL_NAME "CRisprASSembler"
#define PACKAGE_MAJOR_VERSION 1
#define PACKAGE_MINOR_VERSION 0
#define PACKAGE_REVISION 1
#define HAVE_ZLIB 1
#endif
