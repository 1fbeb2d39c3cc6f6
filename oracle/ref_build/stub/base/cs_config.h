/* Hand-written stand-in for the autoconf-generated base/cs_config.h of the
 * reference (code_saturne).  TEST INFRASTRUCTURE ONLY: lets oracle/Makefile.ref
 * compile the reference sources in place (serial: no MPI, OpenMP on) so the
 * real cs_sles / cs_multigrid CPU path can pin the oracle.  Not product code. */
#ifndef CSB200_REF_STUB_CS_CONFIG_H
#define CSB200_REF_STUB_CS_CONFIG_H

#define PACKAGE "code_saturne"
#define PACKAGE_NAME "code_saturne"
#define PACKAGE_TARNAME "code_saturne"
#define PACKAGE_VERSION "9.2-alpha"
#define PACKAGE_STRING "code_saturne 9.2-alpha"
#define PACKAGE_BUGREPORT "none"
#define PACKAGE_URL "none"
#define VERSION "9.2-alpha"
#define LOCALEDIR "/nonexistent/locale"
#define PKGDATADIR "/nonexistent/share"
#define PKGLIBDIR "/nonexistent/lib"

#define HAVE_STDINT_H 1
#define HAVE_STDDEF_H 1
#define HAVE_INTTYPES_H 1
#define HAVE_INT32_T 1
#define HAVE_INT64_T 1
#define HAVE_UINT32_T 1
#define HAVE_UINT64_T 1
#define HAVE_LONG_LONG 1
#define HAVE_LONG_GNUM 1
#define HAVE_PTRDIFF_T 1
#define HAVE_RESTRICT 1
#define HAVE_OPENMP 1
#define HAVE_INTERNAL_PLE 1

#define SIZEOF_SHORT 2
#define SIZEOF_INT 4
#define SIZEOF_LONG 8
#define SIZEOF_LONG_LONG 8
#define SIZEOF_FLOAT 4
#define SIZEOF_DOUBLE 8
#define SIZEOF_VOID_P 8

#define HAVE_UNISTD_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_SYS_STAT_H 1
#define HAVE_SYS_RESOURCE_H 1
#define HAVE_SYS_UTSNAME_H 1
#define HAVE_SYS_SYSINFO_H 1
#define HAVE_DIRENT_H 1
#define HAVE_FCNTL_H 1
#define HAVE_GETCWD 1
#define HAVE_MKDIR 1
#define HAVE_ACCESS 1
#define HAVE_GETRUSAGE 1
#define HAVE_GETTIMEOFDAY 1
#define HAVE_CLOCK_GETTIME 1
#define HAVE_SNPRINTF 1
#define HAVE_SLEEP 1
#define HAVE_SIGACTION 1
#define HAVE_UNAME 1
#define HAVE_SYSINFO 1
#define HAVE_POSIX_MEMALIGN 1
#define HAVE_FSEEKO 1
#define HAVE_GETPWUID 1
#define HAVE_GETEUID 1
#define HAVE_DUP2 1
#define HAVE_SOCKET 1
#define HAVE_DLOPEN 1

#endif
