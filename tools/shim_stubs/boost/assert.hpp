#pragma once
#include <cassert>
#ifndef BOOST_ASSERT
#define BOOST_ASSERT(x) assert(x)
#endif
#define BOOST_VERIFY(x) ((void)(x))
