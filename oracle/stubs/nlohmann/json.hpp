// TEST INFRASTRUCTURE ONLY.  The reference's io_utils.hpp includes <nlohmann/json.hpp> (un-vendored, v3.11.2) and
// never uses it; an empty stand-in lets io_utils.cpp compile in place.
#pragma once
