// Types.h -- working precision and index types (reference: src/Types/Types.h:7-12).
#ifndef PHASE_B200_TYPES_H
#define PHASE_B200_TYPES_H
#include <cstddef>
#include <functional>

typedef double Scalar;
typedef std::size_t Label;
typedef std::size_t Size;
typedef int Index;

template <class T> using Ref = std::reference_wrapper<T>;
#endif
