// TEST INFRASTRUCTURE ONLY: the reference's equalizer.cpp:131 only asks boost::filesystem whether
// its cache file exists; in the oracle build nothing is ever cached.
#pragma once
#include <string>
namespace boost { namespace filesystem {
inline bool exists(const std::wstring &) { return false; }
}}
