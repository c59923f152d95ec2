/*
 * free_interpose.c -- optional, zero-patch reclamation of device-resident results.
 *
 * The unchanged plumbing releases a handle's old value with a plain free(payload)
 * (/root/reference/src/client_context.c:35,82).  Linking this file into the server
 * executable makes that free() tell the shim first, so the HBM buffer registered under the
 * payload address goes back to the engine's pool at the same moment.  Hosts that can take
 * a two-line patch call adb_host_result_release() instead and leave this file out
 * (INTEGRATION.md).
 *
 * A definition of free() in the executable pre-empts libc's for the whole process; every
 * call is forwarded to glibc's real entry point, __libc_free.
 */
#include <stddef.h>

#include "adb_query_api.h"

extern void __libc_free(void *ptr);

void free(void *ptr) {
    if (ptr) adb_host_payload_freed(ptr);      /* one load + compare when nothing is registered */
    __libc_free(ptr);
}
