"""Drop-in mirror of the reference's ``environment`` package: same module / class / method names and signatures
(SURVEY.md section 8-b), executed by the CUDA backend of antsrl_b200.  Put ``antsrl_b200.dropin_path()`` on
sys.path instead of the reference checkout."""
