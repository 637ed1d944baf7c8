// iofunc_shim.cpp -- readStdinBlockData (reference include/iofunc.h:28,
// src/iofunc.cpp:62-69) on top of fmrx_u8_to_f32.  Like the reference it reads
// num_samples bytes from std::cin into a scratch buffer and fills the caller's
// pre-sized vector; on a short read std::cin's state is left failed for the caller's
// rdstate() check (src/project.cpp:51) and the conversion still runs over the buffer.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "fmrx.h"
#include "fmrx_filter.hpp"

void readStdinBlockData(unsigned int num_samples, unsigned int /*block_id*/, std::vector<float> &block_data)
{
    std::vector<uint8_t> raw(num_samples, 0);
    std::cin.read(reinterpret_cast<char *>(raw.data()), num_samples);
    if (std::cin.rdstate() != 0)
        return;                                  // EOF: the caller exits; nothing to convert
    const int rc = fmrx_u8_to_f32(raw.data(), num_samples, block_data.data());
    if (rc != FMRX_OK) {
        std::fprintf(stderr, "readStdinBlockData: %s (%s)\n", fmrx_strerror(rc), fmrx_last_error());
        std::exit(1);
    }
}
