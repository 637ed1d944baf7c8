// project_main.cpp -- the `project` command line on the B200 pipeline.
//
//   project                      mode 0 (message says "mono", like the reference)
//   project <mode 0-3> <1|2|m|s> [--taps N] [--chunk-blocks K] [--device D]
//
// stdin : raw interleaved u8 I/Q at the mode's RF rate
// stdout: raw s16le PCM, interleaved R,L, 48 or 44.1 kS/s -- for either channel
//         argument, exactly as the reference (src/project.cpp:301-302,183-193: the
//         argument only changes the stderr message)
// stderr: "Operating in mode M, mono|stereo", then "End of input stream reached!" at
//         EOF; exit status 1 (src/project.cpp:51-54).  A trailing partial block is
//         dropped as the reference drops it; unlike the reference (which loses up to 4
//         queued blocks when rf_thread calls exit) every complete block is emitted.
//
// Replaces the reference's rf_thread / audio_thread pair and their queue
// (src/project.cpp:19-197): stdin is read K blocks at a time into pinned memory and
// handed to fmrx_process(), whose three CUDA streams overlap copy-in, the four
// kernels and copy-out.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>

#include "fmrx.h"

namespace {

[[noreturn]] void usage(const char *argv0)
{
    std::fprintf(stderr, "Usage: %s\nor \nUsage %s <mode> <channels>\n"
                         "\t\t<mode> is a value from 0 to 3\n\t\t   <channels> is either 1 or 2\n",
                 argv0, argv0);
    std::exit(1);
}

size_t read_full(int fd, uint8_t *dst, size_t want)
{
    size_t got = 0;
    while (got < want) {
        const ssize_t r = ::read(fd, dst + got, want - got);
        if (r <= 0)
            break;
        got += static_cast<size_t>(r);
    }
    return got;
}

bool write_full(int fd, const uint8_t *src, size_t n)
{
    while (n) {
        const ssize_t w = ::write(fd, src, n);
        if (w <= 0)
            return false;
        src += w;
        n -= static_cast<size_t>(w);
    }
    return true;
}

}  // namespace

int main(int argc, char *argv[])
{
    int mode = 0, channels = 1, taps = 51, chunk_blocks = 0, device = -1;
    // positional part, as src/project.cpp:278-299
    int npos = 0;
    const char *pos[2] = { nullptr, nullptr };
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&](int &dst) {
            if (i + 1 >= argc) usage(argv[0]);
            dst = std::atoi(argv[++i]);
        };
        if (a == "--taps") next(taps);
        else if (a == "--chunk-blocks") next(chunk_blocks);
        else if (a == "--device") next(device);
        else if (npos < 2) pos[npos++] = argv[i];
        else usage(argv[0]);
    }
    if (npos < 2) {
        std::fprintf(stderr, "Operating in default mode 0, mono\n");      // a single argument is ignored
    } else {
        mode = std::atoi(pos[0]);
        if (!std::strcmp(pos[1], "m")) channels = 1;
        else if (!std::strcmp(pos[1], "s")) channels = 2;
        else channels = std::atoi(pos[1]);
        if (mode < 0 || mode > 3) {
            std::fprintf(stderr, "Invalid mode: %d!\n", mode);
            return 1;
        }
        if (channels < 1 || channels > 2) {
            std::fprintf(stderr, "Invaild channel: %d!\n", channels);      // sic, as the reference
            return 1;
        }
    }
    std::fprintf(stderr, "Operating in mode %d, %s\n", mode, channels == 1 ? "mono" : "stereo");

    fmrx_mode_info mi;
    if (fmrx_mode_table(mode, taps, &mi) != FMRX_OK) {
        std::fprintf(stderr, "Invalid taps: %d!\n", taps);
        return 1;
    }
    if (chunk_blocks <= 0) {
        // about 0.2 s of signal per call: low latency for a live rtl_sdr pipe
        chunk_blocks = static_cast<int>(0.2 * mi.rf_fs * 2 / mi.block_size);
        if (chunk_blocks < 1) chunk_blocks = 1;
    }
    fmrx_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.mode = mode;
    cfg.taps = taps;
    cfg.n_captures = 1;
    cfg.device = device;
    cfg.chunk_blocks = chunk_blocks;
    fmrx_pipeline *p = nullptr;
    int rc = fmrx_create(&p, &cfg);
    if (rc != FMRX_OK) {
        std::fprintf(stderr, "fmrx_create: %s (%s)\n", fmrx_strerror(rc), fmrx_last_error());
        return 1;
    }
    const size_t in_bytes = static_cast<size_t>(chunk_blocks) * mi.block_size;
    const size_t out_elems = static_cast<size_t>(chunk_blocks) * 2 * mi.audio_per_block;
    void *iq = nullptr, *pcm = nullptr;
    if (fmrx_host_alloc(&iq, in_bytes) != FMRX_OK || fmrx_host_alloc(&pcm, out_elems * sizeof(int16_t)) != FMRX_OK) {
        std::fprintf(stderr, "pinned allocation failed (%s)\n", fmrx_last_error());
        return 1;
    }
    for (;;) {
        const size_t got = read_full(STDIN_FILENO, static_cast<uint8_t *>(iq), in_bytes);
        const size_t blocks = got / mi.block_size;                  // partial block: dropped
        if (blocks) {
            rc = fmrx_process(p, static_cast<const uint8_t *>(iq), in_bytes, blocks,
                              static_cast<int16_t *>(pcm), out_elems);
            if (rc != FMRX_OK) {
                std::fprintf(stderr, "fmrx_process: %s (%s)\n", fmrx_strerror(rc), fmrx_last_error());
                return 1;
            }
            if (!write_full(STDOUT_FILENO, static_cast<const uint8_t *>(pcm),
                            blocks * 2 * mi.audio_per_block * sizeof(int16_t)))
                return 1;
        }
        if (got < in_bytes)
            break;
    }
    std::fprintf(stderr, "End of input stream reached!\n");
    fmrx_host_free(iq);
    fmrx_host_free(pcm);
    fmrx_destroy(p);
    return 1;                                                       // the reference's exit status at EOF
}
