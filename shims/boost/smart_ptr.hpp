// stand-in for boost/smart_ptr.hpp: only what the vendored soth (external/soth-ext) uses; TEST INFRASTRUCTURE (oracle build)
#pragma once
#include <memory>
namespace boost { using std::shared_ptr; using std::make_shared; }
