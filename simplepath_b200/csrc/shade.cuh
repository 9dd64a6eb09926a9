// placeholder — filled in with the shading model
#pragma once
