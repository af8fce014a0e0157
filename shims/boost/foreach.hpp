// stand-in for boost/foreach.hpp: only what the vendored soth (external/soth-ext) uses; TEST INFRASTRUCTURE (oracle build)
#pragma once
#define BOOST_FOREACH(decl, range) for (decl : range)
