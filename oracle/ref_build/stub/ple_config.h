/* Hand-written stand-in for libple's generated ple_config.h (oracle build only). */
#ifndef __PLE_CONFIG_H__
#define __PLE_CONFIG_H__
#define PLE_VERSION "2.0.5"
#define PLE_MAJOR_VERSION 2
#define PLE_MINOR_VERSION 0
#define PLE_RELEASE_VERSION 5
#undef PLE_HAVE_MPI
#endif
