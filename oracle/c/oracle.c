/* Oracle C restatement of the reference overlap aligner olc.nw.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): linked by tests, smoke()
 * and bench.py's cpu_baseline leg, never by the product library.
 *
 * Follows /root/reference/olc.py:40-107 literally (full score and pointer
 * tables, then the do-while traceback).  See oracle/nw_py.py for the list of
 * behaviours and their line citations; tests/test_oracle_nw.py checks that this
 * function, the pure-Python restatement and the golden vectors produced by the
 * reference itself agree.
 */
#include <stdlib.h>
#include <string.h>

/* out[0]=prej out[1]=j_start out[2]=prei out[3]=i_start out[4]=max_i
 * out[5]=alignment length.  a1/a2 need room for m+n+1 bytes. */
int oracle_nw(const char *s1, int m, const char *s2, int n, int *out, char *a1, char *a2)
{
    if (m <= 0 || n <= 0) return -2;
    size_t w = (size_t)m + 1;
    int *score = (int *)calloc((size_t)(n + 1) * w, sizeof(int));
    unsigned char *ptr = (unsigned char *)malloc((size_t)(n + 1) * w);
    if (!score || !ptr) { free(score); free(ptr); return -1; }
    for (int i = 0; i <= n; i++) ptr[(size_t)i * w] = 1;      /* olc.py:56-57 */
    for (int j = 0; j <= m; j++) ptr[j] = 2;                    /* olc.py:58-59 */
    for (int i = 1; i <= n; i++) {                              /* olc.py:62-74 */
        const int *prev = score + (size_t)(i - 1) * w;
        int *cur = score + (size_t)i * w;
        unsigned char *prow = ptr + (size_t)i * w;
        char b = s2[i - 1];
        for (int j = 1; j <= m; j++) {
            int diag = prev[j - 1] + (s1[j - 1] == b ? 1 : -2);
            int up = cur[j - 1] - 2;
            int left = prev[j] - 2;
            int best = left;
            if (up > best) best = up;
            if (diag > best) best = diag;
            cur[j] = best;
            prow[j] = (best == diag) ? 3 : (best == up) ? 2 : 1;
        }
    }
    int max_i = -200, i = 0;                                    /* olc.py:79-83 */
    for (int ii = 0; ii <= n; ii++) {
        int v = score[(size_t)ii * w + m];
        if (v >= max_i) { max_i = v; i = ii; }
    }
    int j = m;
    out[0] = j; out[2] = i; out[4] = max_i;
    int len = 0;
    char *t1 = a1 + (m + n + 1), *t2 = a2 + (m + n + 1);       /* build backwards */
    for (;;) {                                                   /* olc.py:90-105 */
        unsigned char p = ptr[(size_t)i * w + j];
        if (p == 3)      { *--t1 = s1[j - 1]; *--t2 = s2[i - 1]; i--; j--; }
        else if (p == 2) { *--t1 = s1[j - 1]; *--t2 = '-'; j--; }
        else             { *--t1 = '-'; *--t2 = s2[i - 1]; i--; }
        len++;
        if (i == 0 || j == 0) break;
    }
    memmove(a1, t1, (size_t)len);
    memmove(a2, t2, (size_t)len);
    out[1] = j; out[3] = i; out[5] = len;
    free(score); free(ptr);
    return 0;
}
