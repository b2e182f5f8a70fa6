// TEST INFRASTRUCTURE ONLY: stand-in for MSVC <intrin.h> (oracle/_ref build); the reference's
// timestamp.h:21 wants __rdtsc (its only user is the dead delay.cpp).
#pragma once
#include <x86intrin.h>
