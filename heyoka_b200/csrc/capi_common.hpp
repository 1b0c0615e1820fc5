// Shared between the C ABI translation units: error plumbing.
#ifndef HEYOKA_B200_CSRC_CAPI_COMMON_HPP
#define HEYOKA_B200_CSRC_CAPI_COMMON_HPP

#include <stdexcept>
#include <charconv>
#include <string>

namespace heyoka_b200::detail
{

// Shortest round-trip decimal representation of a double ("1", "0.25", "1e+30"), like fmt's "{}".
inline std::string fmt_double(double x)
{
    char buf[64];
    const auto res = std::to_chars(buf, buf + sizeof(buf), x);
    return std::string(buf, res.ptr);
}


struct cuda_error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

void set_last_error(const std::string &);
// To be called inside a catch (...) block: stores the message, returns the HY_ERR_* code.
int translate_exception();

} // namespace heyoka_b200::detail

#endif
