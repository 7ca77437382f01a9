// stand-in: aicp_classification/abstract_classification.hpp includes OpenCV's core header but uses nothing from it
#pragma once
