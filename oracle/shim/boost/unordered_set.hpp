// TEST INFRASTRUCTURE — minimal stand-in for <boost/unordered_set.hpp>.
#pragma once
#include <unordered_set>
namespace boost {
template <typename T>
using unordered_set = std::unordered_set<T>;
}
