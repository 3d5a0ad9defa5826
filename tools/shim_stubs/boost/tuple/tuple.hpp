#pragma once
#include <tuple>
namespace boost { using std::tuple; using std::make_tuple; using std::get; using std::tie; }
