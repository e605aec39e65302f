#pragma once
#include <cstdio>
#define PCL_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
#define PCL_WARN(...) std::fprintf(stderr, __VA_ARGS__)
#define PCL_DEBUG(...) do {} while (0)
