// stand-in for boost/version.hpp: only what the vendored soth (external/soth-ext) uses; TEST INFRASTRUCTURE (oracle build)
#pragma once
#define BOOST_VERSION 107400
