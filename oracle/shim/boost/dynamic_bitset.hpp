// TEST INFRASTRUCTURE — <boost/dynamic_bitset.hpp> is included by maximum_clique.h:42 but never used.
#pragma once
