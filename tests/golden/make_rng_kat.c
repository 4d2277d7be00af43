/* Known-answer vectors of the counter sample stream (include/alvrl_rng.h): gcc -I include tests/golden/make_rng_kat.c -o /tmp/k && /tmp/k > tests/golden/rng_kat.txt */
#include <stdio.h>
#include "alvrl_rng.h"
int main(void) {
    const uint64_t seeds[3] = {0ull, 1234ull, 0xdeadbeefcafef00dull};
    for (int s = 0; s < 3; s++)
        for (uint32_t a = 0; a < 3; a++) {
            uint32_t key = alvrl_rng_key(seeds[s], ALVRL_RNG_CLUSTER, a * 37u, 0);
            uint32_t nk = alvrl_rng_node_key(key, a * 1000u, a * 1000u + 17u + a);
            printf("%llu %u %08x %08x %08x %.9g %.9g\n", (unsigned long long) seeds[s], a, key, nk, alvrl_rng_bits(nk, 1),
                   (double) alvrl_rng_uniform(key, 5), (double) alvrl_rng_uniform(nk, 0));
        }
    return 0;
}
