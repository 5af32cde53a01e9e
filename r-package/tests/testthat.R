library(testthat)
library(nullrangesB200)
test_check("nullrangesB200")
