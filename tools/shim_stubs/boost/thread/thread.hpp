#pragma once
#include <boost/thread/mutex.hpp>
#include <thread>
namespace boost { using std::thread; namespace this_thread { using namespace std::this_thread; } }
