// TEST INFRASTRUCTURE — minimal stand-in for <boost/foreach.hpp> so the reference's src/common compiles here.
#pragma once
#define BOOST_FOREACH(decl, container) for (decl : container)
