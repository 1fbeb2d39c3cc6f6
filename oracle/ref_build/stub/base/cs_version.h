/* Hand-written stand-in for the generated base/cs_version.h (oracle build only). */
#ifndef __CS_VERSION_H__
#define __CS_VERSION_H__
#define CS_VERSION  "9.2-alpha"
#define CS_REVISION "47d979af"
#endif
