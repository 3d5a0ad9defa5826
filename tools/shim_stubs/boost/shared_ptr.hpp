// Minimal stand-ins for the Boost headers the OpenRAVE public headers include, mapped onto the C++17 standard library.
// ONLY for `g++ -fsyntax-only` of plugins/cudacollision/ in an image without Boost (tools/shim_syntax_check.sh); a real build
// uses the real Boost the OpenRAVE install was built with.
#pragma once
#include <memory>
#include <boost/assert.hpp>
namespace boost {
using std::shared_ptr;
using std::weak_ptr;
using std::enable_shared_from_this;
using std::make_shared;
using std::static_pointer_cast;
using std::dynamic_pointer_cast;
using std::const_pointer_cast;
using std::unique_ptr;
}
