/* Unit-cost Levenshtein distance (insert = delete = substitute = 1): what Levenshtein.distance of
 * python-Levenshtein documents.  Two-row dynamic program. */
#include <stdlib.h>
#include <string.h>

int lev_distance(const unsigned char *a, int n, const unsigned char *b, int m)
{
    if (n == 0) return m;
    if (m == 0) return n;
    int stack[130];
    int *row = m + 1 <= 130 ? stack : (int *)malloc(sizeof(int) * (size_t)(m + 1));
    if (!row) return -1;
    for (int j = 0; j <= m; ++j) row[j] = j;
    for (int i = 1; i <= n; ++i) {
        int diag = row[0];
        row[0] = i;
        for (int j = 1; j <= m; ++j) {
            const int up = row[j];
            int v = diag + (a[i - 1] != b[j - 1]);
            if (up + 1 < v) v = up + 1;
            if (row[j - 1] + 1 < v) v = row[j - 1] + 1;
            diag = up;
            row[j] = v;
        }
    }
    const int d = row[m];
    if (row != stack) free(row);
    return d;
}
