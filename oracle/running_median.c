/* CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C restatement of the reference's native running median:
 *   ref: decision_tree/np_stream_median.pyx:8-39 (cummedian)
 * which keeps a max-heap of the lower half and a min-heap of the upper half
 * (the reference stores the upper half negated in a second max-heap,
 * pyx:21,24,27; the values that come out are the same).
 *
 * out[i*2+0], out[i*2+1] = (lower, upper) median of a[0..i].
 * For a prefix of length m these are the order statistics floor((m-1)/2)
 * and floor(m/2): pure selection, no arithmetic on the values.
 */
#include <stdlib.h>

typedef struct {
    double *v;
    long n;
    int is_max; /* 1: largest on top, 0: smallest on top */
} heap_t;

static int above(const heap_t *h, double a, double b) {
    return h->is_max ? (a > b) : (a < b);
}

static void heap_push(heap_t *h, double x) {
    long i = h->n++;
    while (i > 0) {
        long p = (i - 1) / 2;
        if (!above(h, x, h->v[p])) break;
        h->v[i] = h->v[p];
        i = p;
    }
    h->v[i] = x;
}

static double heap_pop(heap_t *h) {
    double top = h->v[0];
    double x = h->v[--h->n];
    long i = 0;
    for (;;) {
        long c = 2 * i + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && above(h, h->v[c + 1], h->v[c])) c++;
        if (!above(h, h->v[c], x)) break;
        h->v[i] = h->v[c];
        i = c;
    }
    if (h->n > 0) h->v[i] = x;
    return top;
}

int oracle_running_median_pairs(const double *a, long n, long stride, double *out) {
    if (n <= 0) return 0;
    heap_t low = {malloc(sizeof(double) * (size_t)(n + 1)), 0, 1};
    heap_t high = {malloc(sizeof(double) * (size_t)(n + 1)), 0, 0};
    if (!low.v || !high.v) {
        free(low.v);
        free(high.v);
        return 1;
    }
    /* pyx:14-16: the first element seeds the lower heap */
    heap_push(&low, a[0]);
    out[0] = out[1] = a[0];
    for (long i = 1; i < n; i++) {
        double x = a[i * stride];
        /* pyx:18-21: route by comparison with the lower half's maximum */
        if (x <= low.v[0])
            heap_push(&low, x);
        else
            heap_push(&high, x);
        /* pyx:23-28: keep the sizes within one of each other */
        if (low.n > high.n + 1)
            heap_push(&high, heap_pop(&low));
        else if (high.n > low.n + 1)
            heap_push(&low, heap_pop(&high));
        /* pyx:30-38 */
        if (low.n == high.n) {
            out[2 * i] = low.v[0];
            out[2 * i + 1] = high.v[0];
        } else if (low.n > high.n) {
            out[2 * i] = out[2 * i + 1] = low.v[0];
        } else {
            out[2 * i] = out[2 * i + 1] = high.v[0];
        }
    }
    free(low.v);
    free(high.v);
    return 0;
}
