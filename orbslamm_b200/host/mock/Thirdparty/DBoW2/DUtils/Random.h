// Stand-in for Thirdparty/DBoW2/DUtils/Random.h (RandomInt as in Random.cpp:47-50).  Test scaffolding only.
#pragma once
#include <cstdlib>
namespace DUtils
{
class Random
{
public:
    static int RandomInt(int min, int max) { int d = max - min + 1; return int(((double)rand() / ((double)RAND_MAX + 1.0)) * d) + min; }
};
}  // namespace DUtils
