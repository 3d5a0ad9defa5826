#pragma once
#include <boost/bind/bind.hpp>
using namespace boost::placeholders;
