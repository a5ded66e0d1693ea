#pragma once
#include "Header.h"
namespace std_msgs { struct String { std::string data; }; }
