/* Force-included (-include) before every reference translation unit: the reference relies on MSVC's
 * transitive includes; g++ 13 needs these spelled out.  Adds no declarations. */
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstring>
#include <memory>
#include <numbers>
#include <optional>
#include <ranges>
#include <string>
#include <string_view>
#include <utility>
#include <vector>
