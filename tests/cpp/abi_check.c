/* A C99 program against include/kopek_cuda.h: every multi-GPU entry point as INTEGRATION.md section 5a shows them.
 * Built by tests/test_abi.py (gcc -std=c99 -Wall -Wextra); without a GPU it must fail loudly -- there is no CPU path. */
#include <stdio.h>
#include "kopek_cuda.h"
int main(void) {
    int32_t devs[] = {0, 1};
    kopek_mgpu *m;
    if (kopek_mgpu_create(&m, devs, 2, 4096, 1024, KOPEK_WINDOW_HANN, KOPEK_OUT_MAG, KOPEK_BINS_HALF, 0) != KOPEK_OK) {
        fprintf(stderr, "%s\n", kopek_last_error());
        return 1;
    }
    uint64_t frames = 0, s0, cnt, f0, nf;
    kopek_mgpu_shard(m, 1 << 20, 1, &s0, &cnt, &f0, &nf);
    float *h = 0; void *spec = 0;
    kopek_mgpu_exec_host(m, h, 0, spec, &frames);
    const float *d_in[2] = {0, 0}; uint64_t n_in[2] = {0, 0}; void *d_out[2] = {0, 0}; uint64_t fr[2];
    kopek_mgpu_exec(m, d_in, n_in, d_out, fr);
    kopek_mgpu_gather(m, d_out, fr, 0, KOPEK_GATHER_NCCL);
    kopek_mgpu_exec_gathered(m, d_in, n_in, 0, &frames);
    float ms; kopek_mgpu_sync(m, &ms);
    kopek_mgpu_destroy(m);
    return 0;
}
