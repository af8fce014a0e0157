// stand-in for boost/noncopyable.hpp: only what the vendored soth (external/soth-ext) uses; TEST INFRASTRUCTURE (oracle build)
#pragma once
namespace boost { class noncopyable { protected: noncopyable() = default; ~noncopyable() = default; noncopyable(const noncopyable&) = delete; noncopyable& operator=(const noncopyable&) = delete; }; }
