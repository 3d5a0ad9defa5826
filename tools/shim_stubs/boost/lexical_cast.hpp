#pragma once
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
struct bad_lexical_cast : std::bad_cast {};
template <typename To, typename From>
To lexical_cast(const From& v)
{
    std::stringstream ss;
    To out;
    ss << v;
    if (!(ss >> out)) throw bad_lexical_cast();
    return out;
}
}
