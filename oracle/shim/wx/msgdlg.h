// test-infrastructure shim: lets the reference OneDimAreaFunction.h compile without wxWidgets
#pragma once
#include <string>
typedef std::string wxString;
