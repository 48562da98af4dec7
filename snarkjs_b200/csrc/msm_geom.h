#pragma once
#include <cstdint>
namespace sb {
struct MsmGeom {
    int c;            // window bits
    int W;            // windows
    uint32_t B;       // buckets per window = 2^(c-1)
};
}
