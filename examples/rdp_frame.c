/* rdp_frame.c -- the smallest C caller of librdp.so: what a testbench-side bridge (a DPI-C / VHPI foreign
 * subprogram of the simulator, or a plain C test driver) would do instead of simulating radar_core.
 *
 *   rdp_frame <n_chirps> <n_samples> <frame.txt> [detections.txt]
 *
 * frame.txt holds one "I Q" pair of int16 per line, chirp-major, sample fast -- the order in which
 * tb_radar_core.stim drives s_axis_tdata (rtl/src/tb_radar_core.vhd:74-123; the same two-column layout as
 * data/golden_input_chirp.txt).  The output is the testbench's own log format: one "r d mag" line per
 * detection (rtl/src/tb_radar_core.vhd:149-160), which model/ADR_visualize.py:109-122 parses.
 *
 * Build:  gcc -std=c99 -O2 -I include examples/rdp_frame.c -L fpga-fmcw-radar-processor_b200 -lrdp \
 *             -Wl,-rpath,'$ORIGIN/../fpga-fmcw-radar-processor_b200' -o examples/rdp_frame
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "rdp.h"

int main(int argc, char** argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s <n_chirps> <n_samples> <frame.txt> [detections.txt]\n", argv[0]);
        return 2;
    }
    const int n_chirps = atoi(argv[1]), n_samples = atoi(argv[2]);
    const long n = (long)n_chirps * n_samples;
    if (n_chirps < 1 || n_samples < 1) { fprintf(stderr, "bad frame size\n"); return 2; }

    int16_t* frame = (int16_t*)malloc(sizeof(int16_t) * 2 * (size_t)n);   /* {I, Q} per sample, I first */
    FILE* in = fopen(argv[3], "r");
    if (!frame || !in) { fprintf(stderr, "cannot read %s\n", argv[3]); return 2; }
    for (long i = 0; i < n; ++i) {
        int re, im;
        if (fscanf(in, "%d %d", &re, &im) != 2) { fprintf(stderr, "%s: %ld samples, expected %ld\n", argv[3], i, n); return 2; }
        frame[2 * i] = (int16_t)re;
        frame[2 * i + 1] = (int16_t)im;
    }
    fclose(in);

    rdp_config cfg;
    rdp_default_config(&cfg);                 /* CFAR 4/4/2/1, rank 75 %, scales 2/4/6 (radar_core.vhd:12-20, :380) */
    cfg.n_chirps = n_chirps;
    cfg.n_samples = n_samples;
    cfg.max_dets_per_cpi = 1 << 16;
    rdp_handle* h = NULL;
    if (rdp_create(&cfg, 0, &h) != RDP_OK) { fprintf(stderr, "rdp_create: %s\n", rdp_last_error(NULL)); return 1; }

    rdp_det* dets = (rdp_det*)malloc(sizeof(rdp_det) * (size_t)cfg.max_dets_per_cpi);
    int32_t n_dets = 0;
    uint32_t status = 0;
    if (!dets || rdp_process_host(h, frame, 1, NULL, dets, &n_dets, &status) != RDP_OK) {
        fprintf(stderr, "rdp_process_host: %s\n", rdp_last_error(h));
        return 1;
    }
    FILE* out = argc > 4 ? fopen(argv[4], "w") : stdout;
    if (!out) { fprintf(stderr, "cannot write %s\n", argv[4]); return 2; }
    const int32_t kept = n_dets < cfg.max_dets_per_cpi ? n_dets : cfg.max_dets_per_cpi;
    for (int32_t i = 0; i < kept; ++i)
        fprintf(out, "%d %d %ld\n", dets[i].range, dets[i].doppler, (long)(dets[i].mag + 0.5f));
    if (out != stdout) fclose(out);
    if (status & RDP_STATUS_DET_OVERFLOW) fprintf(stderr, "detection list overflow: %d found, %d kept\n", n_dets, kept);
    rdp_destroy(h);
    free(dets);
    free(frame);
    return 0;
}
