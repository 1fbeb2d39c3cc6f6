/* Hand-written stand-in for libple's generated ple_config_priv.h (oracle build only). */
#ifndef __PLE_CONFIG_PRIV_H__
#define __PLE_CONFIG_PRIV_H__
#define PACKAGE "ple"
#define HAVE_STDINT_H 1
#define HAVE_GETTIMEOFDAY 1
#define HAVE_GETRUSAGE 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_UNISTD_H 1
#define HAVE_SYS_TIME_H 1
#define HAVE_SYS_RESOURCE_H 1
#endif
