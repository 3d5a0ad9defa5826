#pragma once
#include <functional>
namespace boost {
using std::bind;
using std::ref;
using std::cref;
namespace placeholders { using namespace std::placeholders; }
}
