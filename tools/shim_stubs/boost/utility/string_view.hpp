#pragma once
#include <string_view>
namespace boost { using std::string_view; using std::basic_string_view; }
