"""Step drivers: the loop bodies of the reference's training scripts (SURVEY.md section 3,
row a11 of section 8) built on eadgan_b200.nn / optim."""
