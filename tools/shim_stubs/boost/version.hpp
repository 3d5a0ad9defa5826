#pragma once
#define BOOST_VERSION 107400
#define BOOST_LIB_VERSION "1_74"
