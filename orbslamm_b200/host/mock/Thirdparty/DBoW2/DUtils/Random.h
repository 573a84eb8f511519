// Stand-in for Thirdparty/DBoW2/DUtils/Random.h.  RandomInt has the reference's formula (Random.cpp:47-50) over a private generator instead of
// libc rand(), so that tests are not disturbed by other users of rand() in the process (driver / runtime initialisation).  Test scaffolding only.
#pragma once
#include <cstdlib>
namespace DUtils
{
class Random
{
public:
    static unsigned long long &state() { static unsigned long long s = 88172645463325252ull; return s; }
    static void SeedRand(int seed) { state() = 88172645463325252ull ^ ((unsigned long long)seed * 0x9E3779B97F4A7C15ull); }
    static int next31() { unsigned long long &s = state(); s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (int)((s >> 20) & 0x7fffffff); }
    static int RandomInt(int min, int max) { int d = max - min + 1; return int(((double)next31() / ((double)0x7fffffff + 1.0)) * d) + min; }
};
}  // namespace DUtils
