#pragma once
namespace at { class Tensor; }
