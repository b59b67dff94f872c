"""btAlignedObjectArray::quickSortInternal on island ids (unstable) and the dispatcher's swap-with-last release, as plain Python:
the two operations that decide the solver's manifold order.  Used by the what-if enumerations in this folder."""
import itertools
def bt_quicksort(a, lo, hi):
    i, j = lo, hi
    x = a[(lo + hi) // 2][0]
    while True:
        while a[i][0] < x: i += 1
        while x < a[j][0]: j -= 1
        if i <= j:
            a[i], a[j] = a[j], a[i]; i += 1; j -= 1
        if not (i <= j): break
    if lo < j: bt_quicksort(a, lo, j)
    if i < hi: bt_quicksort(a, i, hi)
def sort(arr, key):
    a = [(key(m), m) for m in arr]
    if len(a) > 1: bt_quicksort(a, 0, len(a) - 1)
    return [m for _, m in a]
def remove(arr, m):
    arr = list(arr); i = arr.index(m); arr[i] = arr[-1]; arr.pop(); return arr
