/* oracle/ref_stubs.c -- TEST INFRASTRUCTURE.  Linker stubs for the five htslib (tabix) entry
 * points that kent/src/lib/linefile.c:240-289 references.  The chain tools never open a tabix
 * file, so these are unreachable; they abort loudly if that assumption ever breaks. */
#include <stdio.h>
#include <stdlib.h>
static void unreachable(const char *name)
{
fprintf(stderr, "oracle/ref_stubs.c: htslib entry %s reached (tabix input is not supported here)\n", name);
abort();
}
void tbx_destroy(void *tbx) { unreachable("tbx_destroy"); }
int hts_itr_next(void *fp, void *iter, void *r, void *data) { unreachable("hts_itr_next"); return -1; }
void hts_itr_destroy(void *iter) { unreachable("hts_itr_destroy"); }
void *hts_get_bgzfp(void *fp) { unreachable("hts_get_bgzfp"); return NULL; }
int hts_close(void *fp) { unreachable("hts_close"); return -1; }
