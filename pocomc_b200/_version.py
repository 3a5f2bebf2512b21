version = "1.2.6+b200.r1"
