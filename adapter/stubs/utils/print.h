// stand-in for ov_core utils/print.h + colors.h
#pragma once
#include <cstdio>
#define PRINT_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
#define RED ""
#define RESET ""
