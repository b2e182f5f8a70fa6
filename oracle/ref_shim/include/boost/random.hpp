// TEST INFRASTRUCTURE ONLY: minimal stand-in so the reference's buffer.hpp (only its declarations
// are needed) parses without Boost. The white-noise generator itself is out of scope.
#pragma once
namespace boost {
struct lagged_fibonacci607 {
    explicit lagged_fibonacci607(unsigned int) {}
    double operator()() { return 0.0; }
};
}
