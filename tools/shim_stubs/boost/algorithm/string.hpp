#pragma once
#include <algorithm>
#include <cctype>
#include <string>
namespace boost {
template <typename A, typename B>
bool iequals(const A& a, const B& b)
{
    const std::string x(a), y(b);
    return x.size() == y.size() && std::equal(x.begin(), x.end(), y.begin(), [](char p, char q) { return std::tolower(p) == std::tolower(q); });
}
template <typename It> struct iterator_range { It b, e; It begin() const { return b; } It end() const { return e; } };
}
