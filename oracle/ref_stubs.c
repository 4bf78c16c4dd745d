/*
 * oracle/ref_stubs.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Link-time stand-ins for the four Xeon-Phi / heterogeneous entry points the
 * reference driver calls for `-m 1` and `-m 2` (declared in the reference at
 * MICsearch.h:29-38 and HETsearch.h:17-26, called from swimm.c:79-118).  Their
 * real bodies use KNC-only intrinsics and Intel LEO offload pragmas and cannot
 * be compiled with gcc, so the oracle binary only supports `-m 0` (the hot
 * path named by BASELINE.json).  Any other mode aborts loudly.
 */
#include <stdio.h>
#include <stdlib.h>

static void unsupported(const char *what)
{
    fprintf(stderr, "oracle/_ref/swimm: %s needs Xeon Phi (KNC) hardware; only -m 0 is built\n", what);
    abort();
}

void mic_search_knc_ap_single_chunk(void)    { unsupported("mic_search_knc_ap_single_chunk"); }
void mic_search_knc_ap_multiple_chunks(void) { unsupported("mic_search_knc_ap_multiple_chunks"); }
void het_search_sse_sp_knc_ap(void)          { unsupported("het_search_sse_sp_knc_ap"); }
void het_search_avx2_sp_knc_ap(void)         { unsupported("het_search_avx2_sp_knc_ap"); }
